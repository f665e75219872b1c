// gmm_kernels.cuh — fused full-covariance EM pass for sm_100a (FP32 CUDA cores).
//
//   gmm_em_full_kernel<D,KP>   one read of z per EM iteration                     (d <= 12)
//       phase 1: one thread per point — Cholesky log-likelihoods for all K components
//                (sklearn _gaussian_mixture.py:490-553), log-sum-exp responsibilities
//                (_base.py:552-582), labels; r_ik parked in shared memory
//       phase 2: one warp per component — the lanes sweep the tile's points and keep the
//                1 + d + d(d+1)/2 moments of "their" component in registers, centred on
//                the current mean (sklearn _gaussian_mixture.py:282-320,168-197)
//
// FP32-FMA-bound, not HBM-bound: ~K(d^2+4d) FMA per point against 4d bytes.
// (A constant-bank + FFMA2 E-step was tried in round 1 and measured 27 % SLOWER at d=9, K=16 —
//  LDCU parameter loads take the same issue slots the shared-memory broadcasts did and the pair
//  accumulators spill under the 128-register cap of a 512-thread CTA)
#pragma once

#include <stdlib.h>

#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

struct GmmArgs {
    const float* z;
    int64_t n;
    int K;
    const float* params;       // mu[K*D], U[K*TRI], cst[K]
    int32_t* labels;
    float* resp;
    const double* ctrl;
    int accumulate;            // 0 E-step only, 1 soft EM, 2 hard (one-hot) responsibilities
    double* stats;
    double* partials;
    unsigned int* counter;
    int skip_reduce;           // leave the per-CTA partial slots for gmm_tail_kernel (fused EM iteration)
    int* grid_out;             // host: number of partial slots written (grid of the statistics kernel)
    unsigned long long* timeline;   // profiling builds (-DSCC_TIMELINE), see scc_common.cuh
};

// after the statistics kernel: stand-alone fixed-order reduction, or hand the slots to the fused tail
static inline int gmm_after_stats(const GmmArgs& a, int grid, int D, cudaStream_t st) {
    if (a.grid_out) *a.grid_out = grid;
    if (a.skip_reduce) return SCC_OK;
    const int NS = SCC_GMM_STAT_DOUBLES(a.K, D);
    reduce_partials_kernel<<<(NS + 255) / 256, 256, 0, st>>>(a.partials, NS, grid, a.stats, a.ctrl);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

__host__ __device__ constexpr int tri(int d) { return d * (d + 1) / 2; }

// ---------------------------------------------------------------------------
// FULL variant
// ---------------------------------------------------------------------------
template <int D, int KP>
__global__ void __launch_bounds__(32 * KP, 1)
gmm_em_full_kernel(const GmmArgs a) {
    constexpr int NT = 32 * KP;
    constexpr int TILE = NT;
    constexpr int S = 3;
    constexpr int TRI = tri(D);
    constexpr int NM = 1 + D + TRI;                      // moments per component
    constexpr int FLUSH = 16;                            // tiles between float -> double flushes
    using Ring = ZRing<D, TILE, S, NT>;
    using L = RowLayout<D>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* r_s = ring_buf + S * Ring::kTileFloats;       // [KP][TILE]
    float* mu_s = r_s + KP * TILE;                       // [KP*D]
    float* u_s = mu_s + ((KP * D + 3) & ~3);             // [KP*TRI]
    float* cst_s = u_s + ((KP * TRI + 3) & ~3);          // [KP]
    double* mom_s = reinterpret_cast<double*>(cst_s + ((KP + 3) & ~3));   // [KP][NM]
    double* ll_s = mom_s + KP * NM;                      // [KP] per-warp log-likelihood
    double* cta_stats = ll_s + KP;                       // [1 + K + K*D + K*TRI]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + 1 + KP * NM);

    if (a.ctrl && a.ctrl[5] != 0.0) return;              // frozen fit: converged or failed earlier

    const int K = a.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < KP * D; i += NT) mu_s[i] = (i < K * D) ? a.params[i] : 0.f;
    for (int i = threadIdx.x; i < KP * TRI; i += NT) u_s[i] = (i < K * TRI) ? a.params[K * D + i] : 0.f;
    if (threadIdx.x < KP) cst_s[threadIdx.x] = ((int)threadIdx.x < K) ? a.params[K * D + K * TRI + threadIdx.x] : 0.f;
    for (int i = threadIdx.x; i < KP * NM; i += NT) mom_s[i] = 0.0;

    Ring ring;
    ring.init(ring_buf, bars, a.z, a.n);
    __syncthreads();
    const int G = gridDim.x;
    for (int s = 0; s < S; ++s) ring.issue(s, blockIdx.x + s * G);
    __syncthreads();

    // phase-2 state of this warp's component
    const int kc = warp;
    float muk[D];
#pragma unroll
    for (int c = 0; c < D; ++c) muk[c] = mu_s[kc * D + c];
    float mom[NM];
#pragma unroll
    for (int s = 0; s < NM; ++s) mom[s] = 0.f;
    float loglik = 0.f;

    auto flush = [&]() {
        if (kc < K) {
#pragma unroll
            for (int s = 0; s < NM; ++s) {
                const float w = warp_sum(mom[s]);
                if (lane == 0) mom_s[kc * NM + s] += (double)w;
                mom[s] = 0.f;
            }
        }
    };

    int it = 0;
    for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G, ++it) {
        const int stage = it % S;
        ring.wait(stage, tile, (uint32_t)(it / S));
        const int np = ring.points(tile);
        const float* ztile = ring.stage_ptr(stage);
        // ---------------- phase 1: E-step for point threadIdx.x ----------------
        {
            const bool active = (int)threadIdx.x < np;
            float lp[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) lp[k] = 0.f;
            float lse = 0.f;
            int label = 0;
            if (active) {
                float x[D];
                load_row<D>(ztile, threadIdx.x, x);
                float best = -3.4e38f;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    lp[k] = -3.4e38f;
                    if (k < K) {
                        float df[D];
#pragma unroll
                        for (int c = 0; c < D; ++c) df[c] = x[c] - mu_s[k * D + c];
                        float m = 0.f;
#pragma unroll
                        for (int b = 0; b < D; ++b) {
                            float y = 0.f;
#pragma unroll
                            for (int c = 0; c <= b; ++c) y = fmaf(df[c], u_s[k * TRI + tri(b) + c], y);
                            m = fmaf(y, y, m);
                        }
                        lp[k] = fmaf(-0.5f, m, cst_s[k]);
                        if (lp[k] > best) { best = lp[k]; label = k; }
                    }
                }
                float se = 0.f;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < K) se += expf(lp[k] - best);
                lse = best + logf(se);
                loglik += lse;
            }
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                float r = (active && k < K) ? expf(lp[k] - lse) : 0.f;
                if ((a.accumulate & 3) == SCC_GMM_HARD) r = (active && k == label) ? 1.f : 0.f;
                r_s[k * TILE + threadIdx.x] = r;
                lp[k] = r;
            }
            if (active) {
                const size_t i = (size_t)tile * TILE + threadIdx.x;
                if (a.labels) a.labels[i] = label;
                if (a.resp) {
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < K) a.resp[i * K + k] = lp[k];
                }
            }
        }
        __syncthreads();
        // ---------------- phase 2: moments of component kc over the tile ----------------
        if ((a.accumulate & 3) && kc < K) {
            for (int t = lane; t < np; t += 32) {
                const float r = r_s[kc * TILE + t];
                float df[D];
                load_row<D>(ztile, t, df);
#pragma unroll
                for (int c = 0; c < D; ++c) df[c] -= muk[c];
                mom[0] += r;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const float w = r * df[c];
                    mom[1 + c] += w;
#pragma unroll
                    for (int b = c; b < D; ++b)                      // S2[c][b], c <= b, column-packed
                        mom[1 + D + tri(b) + c] = fmaf(w, df[b], mom[1 + D + tri(b) + c]);
                }
            }
            if ((it + 1) % FLUSH == 0) flush();
        }
        __syncthreads();
        ring.issue(stage, tile + S * G);
    }
    if (a.accumulate & 3) flush();
    {
        const float w = warp_sum(loglik);
        if (lane == 0) ll_s[warp] = (double)w;
    }
    __syncthreads();
    // pack CTA statistics with the true K: [ll, N_k[K], S1[K*D], S2[K*TRI]]
    const int NS = 1 + K * NM;
    for (int s = threadIdx.x; s < NS; s += NT) {
        double v;
        if (s == 0) {
            v = 0.0;
            for (int w = 0; w < KP; ++w) v += ll_s[w];
        } else if (s < 1 + K) {
            v = mom_s[(s - 1) * NM];
        } else if (s < 1 + K + K * D) {
            const int o = s - 1 - K, k = o / D, c = o - k * D;
            v = mom_s[k * NM + 1 + c];
        } else {
            const int o = s - 1 - K - K * D, k = o / TRI, e = o - k * TRI;
            v = mom_s[k * NM + 1 + D + e];
        }
        cta_stats[s] = v;
    }
    __syncthreads();
    // per-CTA slot; the host-side launcher follows up with reduce_partials_kernel (fixed order)
    for (int s = threadIdx.x; s < NS; s += NT) a.partials[(size_t)blockIdx.x * NS + s] = cta_stats[s];
}

template <int D, int KP>
constexpr size_t gmm_full_smem() {
    constexpr int NT = 32 * KP, TILE = NT, S = 3, TRI = tri(D), NM = 1 + D + TRI;
    return sizeof(float) * (S * TILE * RowLayout<D>::LD + KP * TILE + ((KP * D + 3) & ~3) + ((KP * TRI + 3) & ~3) +
                            ((KP + 3) & ~3)) +
           sizeof(double) * (KP * NM + KP + 1 + KP * NM) + sizeof(uint64_t) * S;
}

template <int D, int KP>
static int launch_gmm_full(const GmmArgs& a, cudaStream_t st) {
    constexpr int NT = 32 * KP;
    auto kern = gmm_em_full_kernel<D, KP>;
    const size_t smem = gmm_full_smem<D, KP>();
    const int64_t tiles = (a.n + NT - 1) / NT;
    int64_t grid = persistent_grid(reinterpret_cast<const void*>(kern), NT, smem, 2);
    if (grid < 0) return (int)grid;
    if (grid > kMaxGmmGrid) grid = kMaxGmmGrid;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, NT, smem, st>>>(a);
    SCC_CUDA(cudaGetLastError());
    return gmm_after_stats(a, (int)grid, D, st);
}


// ---------------------------------------------------------------------------
// PACKED variant of the d <= 12 kernel: same two phases, inner loops in packed FP32 (FFMA2/FADD2).
//   E-step:  y_{2bp}, y_{2bp+1} = sum_c df_c * {U[c][2bp], U[c][2bp+1]}  (U stored as zero-padded
//            pairs in shared memory, df_c splat is free: FFMA2 R, R.F32, R.F32x2, R.F32x2)
//   M-step:  S2[c][2bp..2bp+1] += w_c * {df_2bp, df_2bp+1},  bp >= c/2  (aligned pairs; the one
//            redundant lower-triangle lane of odd rows is discarded at flush time)
// ---------------------------------------------------------------------------
template <int D>
struct GmmPairs {
    static constexpr int DP2 = (D + 1) / 2;
    __host__ __device__ static constexpr int rows(int bp) { return (2 * bp + 2 < D) ? 2 * bp + 2 : D; }
    __host__ __device__ static constexpr int uoff(int bp) { int n = 0; for (int b = 0; b < bp; ++b) n += rows(b); return n; }
    static constexpr int NU2 = uoff(DP2);                       // U pairs per component
    __host__ __device__ static constexpr int moff(int c) { int n = 0; for (int r = 0; r < c; ++r) n += DP2 - r / 2; return n; }
    static constexpr int NP = moff(D);                          // S2 pairs per component
    static constexpr int NPAIR = 1 + DP2 + NP;                  // {S0,-}, S1 pairs, S2 pairs
};

template <int D, int KP>
__global__ void __launch_bounds__(32 * KP, 1)
gmm_em_packed_kernel(const GmmArgs a) {
    constexpr int NT = 32 * KP;
    constexpr int TILE = NT;
    constexpr int S = 3;
    constexpr int TRI = tri(D);
    constexpr int NM = 1 + D + TRI;
    constexpr int FLUSH = 16;
    using Ring = ZRing<D, TILE, S, NT>;
    using P = GmmPairs<D>;
    constexpr int DP2 = P::DP2, NU2 = P::NU2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* r_s = ring_buf + S * Ring::kTileFloats;                       // [KP][TILE]
    float2* nmu2_s = reinterpret_cast<float2*>(r_s + KP * TILE);         // [KP][DP2]   -mu pairs
    float2* u2_s = nmu2_s + ((KP * DP2 + 1) & ~1);                       // [KP][NU2]   U pairs
    float* cst_s = reinterpret_cast<float*>(u2_s + ((KP * NU2 + 1) & ~1));   // [KP]
    double* mom_s = reinterpret_cast<double*>(cst_s + ((KP + 3) & ~3));  // [KP][NM]
    double* ll_s = mom_s + KP * NM;
    double* cta_stats = ll_s + KP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + 1 + KP * NM);

    if (a.ctrl && a.ctrl[5] != 0.0) return;

    const int K = a.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // parameters -> pair layouts
        float* nmu = reinterpret_cast<float*>(nmu2_s);
        for (int i = threadIdx.x; i < KP * DP2 * 2; i += NT) {
            const int k = i / (2 * DP2), c = i - k * (2 * DP2);
            nmu[i] = (k < K && c < D) ? -a.params[k * D + c] : 0.f;
        }
        float* u2 = reinterpret_cast<float*>(u2_s);
        for (int i = threadIdx.x; i < KP * NU2 * 2; i += NT) {
            const int k = i / (2 * NU2), e = i - k * (2 * NU2), pair = e >> 1, ln = e & 1;
            int bp = 0, base = 0;
            while (base + P::rows(bp) <= pair) { base += P::rows(bp); ++bp; }
            const int c = pair - base, b = 2 * bp + ln;
            u2[i] = (k < K && b < D && c <= b) ? a.params[K * D + k * TRI + tri(b) + c] : 0.f;
        }
        if (threadIdx.x < KP) cst_s[threadIdx.x] = ((int)threadIdx.x < K) ? a.params[K * D + K * TRI + threadIdx.x] : 0.f;
        for (int i = threadIdx.x; i < KP * NM; i += NT) mom_s[i] = 0.0;
    }

    Ring ring;
    ring.init(ring_buf, bars, a.z, a.n);
    __syncthreads();
    const int G = gridDim.x;
#pragma unroll
    for (int s = 0; s < S; ++s) ring.issue(s, blockIdx.x + s * G);
    __syncthreads();

    const int kc = warp;
    float2 nmuk[DP2];
#pragma unroll
    for (int c = 0; c < DP2; ++c) nmuk[c] = nmu2_s[kc * DP2 + c];
    float2 mom2[P::NPAIR];
#pragma unroll
    for (int s = 0; s < P::NPAIR; ++s) mom2[s] = make_float2(0.f, 0.f);
    float loglik = 0.f;

    auto flush = [&]() {
        if (kc < K) {
            double* dst = mom_s + kc * NM;
            { const float w = warp_sum(mom2[0].x); if (lane == 0) dst[0] += (double)w; }
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float v = (c & 1) ? mom2[1 + (c >> 1)].y : mom2[1 + (c >> 1)].x;
                const float w = warp_sum(v);
                if (lane == 0) dst[1 + c] += (double)w;
            }
#pragma unroll
            for (int c = 0; c < D; ++c) {
#pragma unroll
                for (int b = c; b < D; ++b) {
                    const int pr = 1 + DP2 + P::moff(c) + (b >> 1) - (c >> 1);
                    const float v = (b & 1) ? mom2[pr].y : mom2[pr].x;
                    const float w = warp_sum(v);
                    if (lane == 0) dst[1 + D + tri(b) + c] += (double)w;
                }
            }
#pragma unroll
            for (int s = 0; s < P::NPAIR; ++s) mom2[s] = make_float2(0.f, 0.f);
        }
    };

    int it = 0;
    for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G, ++it) {
        const int stage = it % S;
        ring.wait(stage, tile, (uint32_t)(it / S));
        const int np = ring.points(tile);
        const float* ztile = ring.stage_ptr(stage);
        // ---------------- phase 1 ----------------
        {
            const bool active = (int)threadIdx.x < np;
            float lp[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) lp[k] = 0.f;
            float lse = 0.f;
            int label = 0;
            if (active) {
                float xr[D];
                load_row<D>(ztile, threadIdx.x, xr);
                float2 x2[DP2];
#pragma unroll
                for (int c = 0; c < DP2; ++c) x2[c] = make_float2(xr[2 * c], (2 * c + 1 < D) ? xr[2 * c + 1] : 0.f);
                float best = -3.4e38f;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    lp[k] = -3.4e38f;
                    if (k < K) {
                        float2 df2[DP2];
#pragma unroll
                        for (int c = 0; c < DP2; ++c) df2[c] = __fadd2_rn(x2[c], nmu2_s[k * DP2 + c]);
                        float2 m2 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int bp = 0; bp < DP2; ++bp) {
                            float2 y2 = make_float2(0.f, 0.f);
#pragma unroll
                            for (int c = 0; c < P::rows(bp); ++c) {
                                const float dc = (c & 1) ? df2[c >> 1].y : df2[c >> 1].x;
                                y2 = __ffma2_rn(make_float2(dc, dc), u2_s[k * NU2 + P::uoff(bp) + c], y2);
                            }
                            m2 = __ffma2_rn(y2, y2, m2);
                        }
                        lp[k] = fmaf(-0.5f, m2.x + m2.y, cst_s[k]);
                        if (lp[k] > best) { best = lp[k]; label = k; }
                    }
                }
                float se = 0.f;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < K) se += expf(lp[k] - best);
                lse = best + logf(se);
                loglik += lse;
            }
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                float r = (active && k < K) ? expf(lp[k] - lse) : 0.f;
                if ((a.accumulate & 3) == SCC_GMM_HARD) r = (active && k == label) ? 1.f : 0.f;
                r_s[k * TILE + threadIdx.x] = r;
                lp[k] = r;
            }
            if (active) {
                const size_t i = (size_t)tile * TILE + threadIdx.x;
                if (a.labels) a.labels[i] = label;
                if (a.resp) {
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < K) a.resp[i * K + k] = lp[k];
                }
            }
        }
        __syncthreads();
        // ---------------- phase 2 ----------------
        if ((a.accumulate & 3) && kc < K) {
            for (int t = lane; t < np; t += 32) {
                const float r = r_s[kc * TILE + t];
                float xr[D];
                load_row<D>(ztile, t, xr);
                float2 df2[DP2], w2[DP2];
                const float2 r2 = splat2(r);
#pragma unroll
                for (int c = 0; c < DP2; ++c) {
                    df2[c] = __fadd2_rn(make_float2(xr[2 * c], (2 * c + 1 < D) ? xr[2 * c + 1] : 0.f), nmuk[c]);
                    w2[c] = __fmul2_rn(r2, df2[c]);
                    mom2[1 + c] = __fadd2_rn(mom2[1 + c], w2[c]);
                }
                mom2[0].x += r;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const float wc = (c & 1) ? w2[c >> 1].y : w2[c >> 1].x;
#pragma unroll
                    for (int bp = c >> 1; bp < DP2; ++bp) {
                        const int pr = 1 + DP2 + P::moff(c) + bp - (c >> 1);
                        mom2[pr] = __ffma2_rn(make_float2(wc, wc), df2[bp], mom2[pr]);
                    }
                }
            }
            if ((it + 1) % FLUSH == 0) flush();
        }
        __syncthreads();
        ring.issue(stage, tile + S * G);
    }
    if (a.accumulate & 3) flush();
    {
        const float w = warp_sum(loglik);
        if (lane == 0) ll_s[warp] = (double)w;
    }
    __syncthreads();
    const int NS = 1 + K * NM;
    for (int s = threadIdx.x; s < NS; s += NT) {
        double v;
        if (s == 0) {
            v = 0.0;
            for (int w = 0; w < KP; ++w) v += ll_s[w];
        } else if (s < 1 + K) {
            v = mom_s[(s - 1) * NM];
        } else if (s < 1 + K + K * D) {
            const int o = s - 1 - K, k = o / D, c = o - k * D;
            v = mom_s[k * NM + 1 + c];
        } else {
            const int o = s - 1 - K - K * D, k = o / TRI, e = o - k * TRI;
            v = mom_s[k * NM + 1 + D + e];
        }
        cta_stats[s] = v;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < NS; s += NT) a.partials[(size_t)blockIdx.x * NS + s] = cta_stats[s];
}

template <int D, int KP>
constexpr size_t gmm_packed_smem() {
    constexpr int NT = 32 * KP, TILE = NT, S = 3, TRI = tri(D), NM = 1 + D + TRI;
    using P = GmmPairs<D>;
    return sizeof(float) * (S * TILE * RowLayout<D>::LD + KP * TILE + 2 * ((KP * P::DP2 + 1) & ~1) + 2 * ((KP * P::NU2 + 1) & ~1) +
                            ((KP + 3) & ~3)) +
           sizeof(double) * (KP * NM + KP + 1 + KP * NM) + sizeof(uint64_t) * S;
}

// ---------------------------------------------------------------------------
// SPARSE variant of the d <= 12 kernel (default).  Same two phases as gmm_em_full_kernel, with the two
// changes that matter for the instruction count (profiles/: the full kernel issues ~3600 thread-instructions
// per point for 1872 algorithmic FMAs at d = 9, K = 16, LSU 37 %):
//   E-step:  packed FP32 over COMPONENT PAIRS (2kp, 2kp+1): df2_c = (x_c - mu_kc, x_c - mu_k'c) with the point's
//            coordinate in the .F32 broadcast operand, y2_b = sum_c df2_c * (U_k[c][b], U_k'[c][b]) — every
//            FADD2/FFMA2 does two components' work and every 128-bit shared-memory load feeds four FMAs; no pad
//            lane for odd d (the pair is over k, not over the dimension).  Responsibilities come from ONE
//            exponential per component (e_k = 2^((lp_k - max) log2 e), r_k = e_k / sum e) instead of two.
//   M-step:  responsibility-sparsity skip (lever c).  After the first few EM iterations only ~1.4 of
//            16 components per point have r_ik >= 2^-30; pairs below that contribute < 1e-9 relative to any
//            moment.  Phase 1 ballots (r_ik >= 2^-30) per component; the warp masks are turned into per-component
//            COMPACT point lists (positions from a prefix over the source warps: point order, deterministic) and
//            the component's warp sweeps only its list, all 32 lanes busy.  SCC_GMM_NOSKIP keeps every pair.
// ---------------------------------------------------------------------------
constexpr float kGmmSkipThreshold = 9.313225746154785e-10f;       // 2^-30

// points per thread of the E-step: every parameter load (one 128-bit shared-memory load per two U pairs) then
// feeds PP points.  With one point per thread the kernel is bound by the shared-memory RETURN bandwidth
// (128 B/clk/SM: a warp-wide LDS.128 occupies the pipe for 4 cycles whatever the broadcast) — K (d + TRI) floats per
// point, 3.4 KB at d = 9, K = 16 = 27 cycles/point against ~8 for the arithmetic.
template <int KP>
__host__ __device__ constexpr int gmm_ppt() { return 2; }
template <int KP>
__host__ __device__ constexpr int gmm_stages() { return KP >= 16 ? 2 : 3; }

template <int D, int KP>
__global__ void __launch_bounds__(32 * KP, 1)
gmm_em_sparse_kernel(const GmmArgs a) {
    constexpr int NT = 32 * KP;
    constexpr int PP = gmm_ppt<KP>();
    constexpr int TILE = NT * PP;
    constexpr int S = gmm_stages<KP>();
    constexpr int NW = KP;                                 // warps per CTA == components
    constexpr int NSRC = NW * PP;                          // 32-point groups of a tile (ballot sources)
    constexpr int JP = KP / 2;                             // component pairs
    constexpr int TRI = tri(D);
    constexpr int TRIP = (TRI + 1) & ~1;                   // U pairs per component pair, padded to 16 bytes
    constexpr int DP = (D + 1) & ~1;
    constexpr int NM = 1 + D + TRI;
    constexpr int FLUSH = 32;                              // tiles between float -> double flushes (<= ~100 terms per lane)
    using Ring = ZRing<D, TILE, S, NT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* r_s = ring_buf + ((S * Ring::kTileFloats + 3) & ~3);            // [KP][TILE]
    float2* nmu2_s = reinterpret_cast<float2*>(r_s + KP * TILE);           // [JP][DP]    -mu pairs over components
    float2* u2_s = nmu2_s + JP * DP;                                       // [JP][TRIP]  U pairs
    float2* cst2_s = u2_s + JP * TRIP;                                     // [JP]
    float* muk_s = reinterpret_cast<float*>(cst2_s + ((JP + 1) & ~1));     // [KP][D] means (phase 2)
    double* mom_s = reinterpret_cast<double*>(muk_s + ((KP * D + 3) & ~3));   // [KP][NM]
    double* ll_s = mom_s + KP * NM;                                        // [KP]
    double* cta_stats = ll_s + KP;                                         // [1 + KP*NM]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + 1 + KP * NM); // [S]
    unsigned int* mask_s = reinterpret_cast<unsigned int*>(bars + S);      // [KP][NSRC] ballot of (r >= thr) per 32-point group
    int* pref_s = reinterpret_cast<int*>(mask_s + KP * NSRC);              // [KP][NSRC + 1] exclusive prefix, total
    unsigned short* list_s = reinterpret_cast<unsigned short*>(pref_s + KP * (NSRC + 1));   // [KP][TILE]

    if (a.ctrl && a.ctrl[5] != 0.0) return;              // frozen fit: converged or failed earlier

    const int K = a.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SCC_TL(a.timeline, 0);
    {   // parameters -> component-pair layouts
        float* nmu = reinterpret_cast<float*>(nmu2_s);
        for (int i = threadIdx.x; i < JP * DP * 2; i += NT) {
            const int kp = i / (2 * DP), e = i - kp * (2 * DP), c = e >> 1, k = 2 * kp + (e & 1);
            nmu[i] = (k < K && c < D) ? -a.params[k * D + c] : 0.f;
        }
        float* u2 = reinterpret_cast<float*>(u2_s);
        for (int i = threadIdx.x; i < JP * TRIP * 2; i += NT) {
            const int kp = i / (2 * TRIP), e = i - kp * (2 * TRIP), t = e >> 1, k = 2 * kp + (e & 1);
            u2[i] = (k < K && t < TRI) ? a.params[K * D + k * TRI + t] : 0.f;
        }
        float* cst = reinterpret_cast<float*>(cst2_s);
        if (threadIdx.x < KP) cst[threadIdx.x] = ((int)threadIdx.x < K) ? a.params[K * D + K * TRI + threadIdx.x] : 0.f;
        for (int i = threadIdx.x; i < KP * D; i += NT) muk_s[i] = (i < K * D) ? a.params[i] : 0.f;
        for (int i = threadIdx.x; i < KP * NM; i += NT) mom_s[i] = 0.0;
    }
    const float thr = (a.accumulate & SCC_GMM_NOSKIP) ? 0.f : kGmmSkipThreshold;
    const int acc_mode = a.accumulate & 3;

    // Blocked partition: CTA b owns a contiguous range of 32-point slices (ranges differ by at most one slice) and
    // walks it in tiles of TILE points with one partial tile at the end — instead of whole tiles dealt out cyclically,
    // where ceil(tiles / grid) rounds of a 12 us tile decide the kernel time (1.25M points per GPU: 8.25 -> 9 rounds).
    const int G = gridDim.x;
    const int64_t slices = (a.n + 31) >> 5;
    const int64_t per = slices / G, rem = slices - per * G;
    const int64_t s0 = (int64_t)blockIdx.x * per + ((int64_t)blockIdx.x < rem ? (int64_t)blockIdx.x : rem);
    const int64_t lo = 32 * s0 < a.n ? 32 * s0 : a.n;
    const int64_t hi_ = 32 * (s0 + per + ((int64_t)blockIdx.x < rem ? 1 : 0));
    const int64_t hi = hi_ < a.n ? hi_ : a.n;
    Ring ring;
    ring.init(ring_buf, bars, a.z + (size_t)lo * D, hi - lo);      // lo is a multiple of 32 points: 16-byte aligned
    __syncthreads();
#pragma unroll
    for (int s = 0; s < S; ++s) ring.issue(s, s);
    __syncthreads();
    SCC_TL(a.timeline, 1);

    // phase-2 state of this warp's component
    const int kc = warp;
    float muk[D];
#pragma unroll
    for (int c = 0; c < D; ++c) muk[c] = muk_s[kc * D + c];
    float mom[NM];
#pragma unroll
    for (int s = 0; s < NM; ++s) mom[s] = 0.f;
    float loglik = 0.f;

    auto flush = [&]() {
        if (kc < K) {
#pragma unroll
            for (int s = 0; s < NM; ++s) {
                const float w = warp_sum(mom[s]);
                if (lane == 0) mom_s[kc * NM + s] += (double)w;
                mom[s] = 0.f;
            }
        }
    };

    int it = 0;
    for (int tile = 0; tile < ring.num_tiles; ++tile, ++it) {
        const int stage = it % S;
        ring.wait(stage, tile, (uint32_t)(it / S));
        const int np = ring.points(tile);
        const float* ztile = ring.stage_ptr(stage);
        if (tile == 0) SCC_TL(a.timeline, 2);
        // ---------------- phase 1: E-step for points threadIdx.x + p * NT ----------------
        bool active[PP];
        float2 r2[PP][JP];
#pragma unroll
        for (int p = 0; p < PP; ++p) {
            active[p] = (int)threadIdx.x + p * NT < np;
#pragma unroll
            for (int kp = 0; kp < JP; ++kp) r2[p][kp] = make_float2(0.f, 0.f);
        }
        if (__any_sync(0xffffffffu, active[0])) {       // warps beyond the end of a partial tile skip the E-step
            float x[PP][D];
#pragma unroll
            for (int p = 0; p < PP; ++p) {
#pragma unroll
                for (int c = 0; c < D; ++c) x[p][c] = 0.f;
                if (active[p]) load_row<D>(ztile, threadIdx.x + p * NT, x[p]);
            }
            float2 lp2[PP][JP];
#pragma unroll
            for (int kp = 0; kp < JP; ++kp) {
#pragma unroll
                for (int p = 0; p < PP; ++p) lp2[p][kp] = make_float2(-3.4e38f, -3.4e38f);
                if (2 * kp < K) {
                    float2 df2[PP][D];
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const float2 nm = nmu2_s[kp * DP + c];
#pragma unroll
                        for (int p = 0; p < PP; ++p) df2[p][c] = __fadd2_rn(make_float2(x[p][c], x[p][c]), nm);
                    }
                    float2 m2[PP];
#pragma unroll
                    for (int p = 0; p < PP; ++p) m2[p] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int b = 0; b < D; ++b) {
                        float2 y2[PP];
#pragma unroll
                        for (int p = 0; p < PP; ++p) y2[p] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int c = 0; c <= b; ++c) {
                            const float2 u = u2_s[kp * TRIP + tri(b) + c];
#pragma unroll
                            for (int p = 0; p < PP; ++p) y2[p] = __ffma2_rn(df2[p][c], u, y2[p]);
                        }
#pragma unroll
                        for (int p = 0; p < PP; ++p) m2[p] = __ffma2_rn(y2[p], y2[p], m2[p]);
                    }
                    const float2 cst = cst2_s[kp];
#pragma unroll
                    for (int p = 0; p < PP; ++p) {
                        float2 v = __ffma2_rn(make_float2(-0.5f, -0.5f), m2[p], cst);
                        if (2 * kp + 1 >= K) v.y = -3.4e38f;
                        lp2[p][kp] = v;
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < PP; ++p) {
                float best = -3.4e38f;
                int label = 0;
#pragma unroll
                for (int kp = 0; kp < JP; ++kp) {
                    if (lp2[p][kp].x > best) { best = lp2[p][kp].x; label = 2 * kp; }
                    if (lp2[p][kp].y > best) { best = lp2[p][kp].y; label = 2 * kp + 1; }
                }
                // e_k = exp(lp_k - best) once per component; r_k = e_k / sum_k e_k; log p(x) = best + log sum_k e_k
                const float2 nb2 = make_float2(-best, -best);
                const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
                float2 se2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int kp = 0; kp < JP; ++kp) {
                    const float2 arg = __fmul2_rn(__fadd2_rn(lp2[p][kp], nb2), l2e);
                    r2[p][kp] = make_float2(ex2_approx(fmaxf(arg.x, -126.f)), ex2_approx(fmaxf(arg.y, -126.f)));
                    if (2 * kp >= K) r2[p][kp].x = 0.f;
                    if (2 * kp + 1 >= K) r2[p][kp].y = 0.f;
                    se2 = __fadd2_rn(se2, r2[p][kp]);
                }
                const float se = se2.x + se2.y;
                if (active[p]) loglik += best + 0.6931471805599453f * lg2_approx(se);
                const float inv = 1.f / se;
#pragma unroll
                for (int kp = 0; kp < JP; ++kp) r2[p][kp] = __fmul2_rn(r2[p][kp], make_float2(inv, inv));
                if (acc_mode == SCC_GMM_HARD) {
#pragma unroll
                    for (int kp = 0; kp < JP; ++kp)
                        r2[p][kp] = make_float2(label == 2 * kp ? 1.f : 0.f, label == 2 * kp + 1 ? 1.f : 0.f);
                }
                if (active[p]) {
                    const size_t i = (size_t)lo + (size_t)tile * TILE + threadIdx.x + p * NT;
                    if (a.labels) a.labels[i] = label;
                    if (a.resp) {
#pragma unroll
                        for (int k = 0; k < KP; ++k)
                            if (k < K) a.resp[i * K + k] = (k & 1) ? r2[p][k / 2].y : r2[p][k / 2].x;
                    }
                }
            }
        }
        if (acc_mode) {
            // ---- per-component ballots of the significant pairs, responsibilities parked for phase 2
#pragma unroll
            for (int p = 0; p < PP; ++p) {
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    const float r = (k & 1) ? r2[p][k / 2].y : r2[p][k / 2].x;
                    r_s[k * TILE + threadIdx.x + p * NT] = r;
                    const unsigned int m = __ballot_sync(0xffffffffu, active[p] && k < K && r >= thr && r > 0.f);
                    if (lane == 0) mask_s[k * NSRC + p * NW + warp] = m;
                }
            }
            __syncthreads();
            // exclusive prefix of the per-group counts: thread (k, group)
            for (int o = threadIdx.x; o < KP * NSRC; o += NT) {
                const int k = o / NSRC, w = o - k * NSRC;
                int pfx = 0;
                for (int ww = 0; ww < w; ++ww) pfx += __popc(mask_s[k * NSRC + ww]);
                pref_s[k * (NSRC + 1) + w] = pfx;
                if (w == NSRC - 1) pref_s[k * (NSRC + 1) + NSRC] = pfx + __popc(mask_s[k * NSRC + w]);
            }
            __syncthreads();
            // compact point lists, point order (deterministic)
            const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
            for (int p = 0; p < PP; ++p) {
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    const unsigned int m = mask_s[k * NSRC + p * NW + warp];
                    if ((m >> lane) & 1u)
                        list_s[k * TILE + pref_s[k * (NSRC + 1) + p * NW + warp] + __popc(m & lt)] =
                            (unsigned short)(threadIdx.x + p * NT);
                }
            }
            __syncthreads();
            // ---------------- phase 2: moments of component kc over its list ----------------
            if (kc < K) {
                const int cnt = pref_s[kc * (NSRC + 1) + NSRC];
                for (int e = lane; e < cnt; e += 32) {
                    const int t = list_s[kc * TILE + e];
                    const float r = r_s[kc * TILE + t];
                    float df[D];
                    load_row<D>(ztile, t, df);
#pragma unroll
                    for (int c = 0; c < D; ++c) df[c] -= muk[c];
                    mom[0] += r;
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const float w = r * df[c];
                        mom[1 + c] += w;
#pragma unroll
                        for (int b = c; b < D; ++b)                      // S2[c][b], c <= b, column-packed
                            mom[1 + D + tri(b) + c] = fmaf(w, df[b], mom[1 + D + tri(b) + c]);
                    }
                }
                if ((it + 1) % FLUSH == 0) flush();
            }
        }
        __syncthreads();
        ring.issue(stage, tile + S);
    }
    pdl_trigger();                      // the fused tail kernel may be scheduled under this kernel's drain
    SCC_TL(a.timeline, 3);
    if (acc_mode) flush();
    {
        const float w = warp_sum(loglik);
        if (lane == 0) ll_s[warp] = (double)w;
    }
    __syncthreads();
    SCC_TL(a.timeline, 4);
    // pack CTA statistics with the true K: [ll, N_k[K], S1[K*D], S2[K*TRI]]
    const int NS = 1 + K * NM;
    for (int s = threadIdx.x; s < NS; s += NT) {
        double v;
        if (s == 0) {
            v = 0.0;
            for (int w = 0; w < KP; ++w) v += ll_s[w];
        } else if (s < 1 + K) {
            v = mom_s[(s - 1) * NM];
        } else if (s < 1 + K + K * D) {
            const int o = s - 1 - K, k = o / D, c = o - k * D;
            v = mom_s[k * NM + 1 + c];
        } else {
            const int o = s - 1 - K - K * D, k = o / TRI, e = o - k * TRI;
            v = mom_s[k * NM + 1 + D + e];
        }
        cta_stats[s] = v;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < NS; s += NT) a.partials[(size_t)blockIdx.x * NS + s] = cta_stats[s];
    SCC_TL(a.timeline, 5);
}

template <int D, int KP>
constexpr size_t gmm_sparse_smem() {
    constexpr int NT = 32 * KP, PP = gmm_ppt<KP>(), TILE = NT * PP, S = gmm_stages<KP>(), TRI = tri(D), NM = 1 + D + TRI;
    constexpr int JP = KP / 2, NSRC = KP * PP;
    constexpr int TRIP = (TRI + 1) & ~1, DP = (D + 1) & ~1;
    return sizeof(float) * (((S * TILE * RowLayout<D>::LD + 3) & ~3) + KP * TILE + 2 * (JP * DP + JP * TRIP + ((JP + 1) & ~1)) +
                            ((KP * D + 3) & ~3)) +
           sizeof(double) * (KP * NM + KP + 1 + KP * NM) + sizeof(uint64_t) * S +
           sizeof(unsigned int) * KP * NSRC + sizeof(int) * KP * (NSRC + 1) + sizeof(unsigned short) * KP * TILE;
}

// which d <= 12 kernel to run (set from measurements; SCC_GMM_FORCE_SCALAR / _PACKED override for A/B runs)
template <int D, int KP>
static int launch_gmm_small(const GmmArgs& a, cudaStream_t st) {
    constexpr int NT = 32 * KP;
    static const int force = []() {
        const char* e = getenv("SCC_GMM_VARIANT");
        return e ? (e[0] == 'p' ? 1 : (e[0] == 's' ? 2 : 0)) : 0;
    }();
    // measured (tools/gmm_ab.py, N=4M): packed wins only for even d (d=12: 477 vs 522 us, d=4: 153 vs 160 us);
    // at the reference's d=9 the pad lane and half-pair splats make it 30 % slower (534 vs 403 us, K=8)
    static const int sparse_off = []() {
        const char* e = getenv("SCC_GMM_VARIANT");
        return e ? (e[0] == 'f' || e[0] == 'p' || e[0] == 's') : 0;          // f(ull) / p(acked) / s(calar): the round-1 kernels
    }();
    if (!sparse_off) {
        auto kern = gmm_em_sparse_kernel<D, KP>;
        constexpr size_t smem = gmm_sparse_smem<D, KP>();
        constexpr int tile_points = NT * gmm_ppt<KP>();
        const int64_t tiles = (a.n + tile_points - 1) / tile_points;       // (no CTA with less than one tile of work)
        int64_t grid = persistent_grid(reinterpret_cast<const void*>(kern), NT, smem, 2);
        if (grid < 0) return (int)grid;
        if (grid > kMaxGmmGrid) grid = kMaxGmmGrid;
        if (grid > tiles) grid = tiles;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, NT, smem, st>>>(a);
        SCC_CUDA(cudaGetLastError());
        return gmm_after_stats(a, (int)grid, D, st);
    }
    const bool packed = force ? (force == 1) : false;
    if (!packed) return launch_gmm_full<D, KP>(a, st);
    auto kern = gmm_em_packed_kernel<D, KP>;
    constexpr size_t smem = gmm_packed_smem<D, KP>();
    const int64_t tiles = (a.n + NT - 1) / NT;
    int64_t grid = persistent_grid(reinterpret_cast<const void*>(kern), NT, smem, 2);
    if (grid < 0) return (int)grid;
    if (grid > kMaxGmmGrid) grid = kMaxGmmGrid;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, NT, smem, st>>>(a);
    SCC_CUDA(cudaGetLastError());
    return gmm_after_stats(a, (int)grid, D, st);
}

// ---------------------------------------------------------------------------
// BLOCK variant (d % 4 == 0, d >= 16): the 1 + d + d(d+1)/2 moments of a component no longer fit
// one thread's registers, so phase 2 tiles the symmetric second moment into 4x4 register blocks:
// a "unit" is (component k, row block ab, column block bb >= ab); the K * NB2 units are dealt to
// the CTA's threads and every thread sweeps ALL points of the tile for its units (r broadcast,
// two LDS.128 of the centred point, 4 FMUL + 16 FMA).  Diagonal units also carry S1 (and unit
// (0,0) S0).  Partials are flushed from fp32 registers straight into the CTA's float64 slot.
// ---------------------------------------------------------------------------
template <int D, int KP>
struct GmmBlock {
    static constexpr int NT = 256, TILE = 256, S = 2;
    static constexpr int NBLK = D / 4;
    static constexpr int NB2 = NBLK * (NBLK + 1) / 2;
    static constexpr int MAXU = (KP * NB2 + NT - 1) / NT;      // units per thread
    static constexpr int TRI = D * (D + 1) / 2;
    static constexpr int NM = 1 + D + TRI;
};

template <int D, int KP>
__global__ void __launch_bounds__(256, 1)
gmm_em_block_kernel(const GmmArgs a) {
    using B = GmmBlock<D, KP>;
    using L = RowLayout<D>;
    constexpr int NT = B::NT, TILE = B::TILE, S = B::S, TRI = B::TRI, NM = B::NM, NBLK = B::NBLK, NB2 = B::NB2;
    constexpr int MAXU = B::MAXU;
    constexpr int FLUSH = 8;
    using Ring = ZRing<D, TILE, S, NT>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* r_s = ring_buf + S * Ring::kTileFloats;       // [TILE][KP]
    float* mu_s = r_s + TILE * KP;                       // [KP*D]
    float* u_s = mu_s + KP * D;                          // [KP*TRI]
    float* cst_s = u_s + ((KP * TRI + 3) & ~3);          // [KP]
    double* ll_s = reinterpret_cast<double*>(cst_s + ((KP + 3) & ~3));   // [NT/32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ll_s + NT / 32);
    constexpr int NWB = NT / 32;
    unsigned int* mask_s = reinterpret_cast<unsigned int*>(bars + S);    // [KP][NWB] ballot of (r >= 2^-30) per warp
    int* pref_s = reinterpret_cast<int*>(mask_s + KP * NWB);             // [KP][NWB + 1] exclusive prefix, total
    unsigned short* list_s = reinterpret_cast<unsigned short*>(pref_s + KP * (NWB + 1));   // [KP][TILE] compact point lists

    if (a.ctrl && a.ctrl[5] != 0.0) return;

    const int K = a.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float thr = (a.accumulate & SCC_GMM_NOSKIP) ? 0.f : kGmmSkipThreshold;
    for (int i = threadIdx.x; i < KP * D; i += NT) mu_s[i] = (i < K * D) ? a.params[i] : 0.f;
    for (int i = threadIdx.x; i < KP * TRI; i += NT) u_s[i] = (i < K * TRI) ? a.params[K * D + i] : 0.f;
    if (threadIdx.x < KP) cst_s[threadIdx.x] = ((int)threadIdx.x < K) ? a.params[K * D + K * TRI + threadIdx.x] : 0.f;
    const int NS = 1 + K * NM;
    double* slot = a.partials + (size_t)blockIdx.x * NS;
    for (int i = threadIdx.x; i < NS; i += NT) slot[i] = 0.0;

    Ring ring;
    ring.init(ring_buf, bars, a.z, a.n);
    __syncthreads();
    const int G = gridDim.x;
#pragma unroll
    for (int s = 0; s < S; ++s) ring.issue(s, blockIdx.x + s * G);
    __syncthreads();

    // this thread's units
    int uk[MAXU], ua[MAXU], ub[MAXU];
    float mua[MAXU][4], mub[MAXU][4];
    float acc[MAXU][16], s1[MAXU][4], s0[MAXU];
    const int nunits = K * NB2;
#pragma unroll
    for (int m = 0; m < MAXU; ++m) {
        const int u = threadIdx.x + m * NT;
        int k = -1, ab = 0, bb = 0;
        if (u < nunits) {
            k = u / NB2;
            int e = u - k * NB2;                       // row-major over the upper block triangle
            while (e >= NBLK - ab) { e -= NBLK - ab; ++ab; }
            bb = ab + e;
        }
        uk[m] = k; ua[m] = ab; ub[m] = bb;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            mua[m][c] = (k >= 0) ? mu_s[k * D + 4 * ab + c] : 0.f;
            mub[m][c] = (k >= 0) ? mu_s[k * D + 4 * bb + c] : 0.f;
            s1[m][c] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[m][e] = 0.f;
        s0[m] = 0.f;
    }
    float loglik = 0.f;

    auto flush = [&]() {
#pragma unroll
        for (int m = 0; m < MAXU; ++m) {
            if (uk[m] >= 0) {
                double* dst = slot + 1;                 // [N_k[K] | S1[K*D] | S2[K*TRI]]
                const int k = uk[m], ab = ua[m], bb = ub[m];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int ra = 4 * ab + r, cb = 4 * bb + c;
                        if (ra <= cb) dst[K + K * D + k * TRI + cb * (cb + 1) / 2 + ra] += (double)acc[m][4 * r + c];
                        acc[m][4 * r + c] = 0.f;
                    }
                if (ab == bb) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) { dst[K + k * D + 4 * ab + r] += (double)s1[m][r]; s1[m][r] = 0.f; }
                    if (ab == 0) { dst[k] += (double)s0[m]; s0[m] = 0.f; }
                }
            }
        }
    };

    int it = 0;
    for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G, ++it) {
        const int stage = it % S;
        ring.wait(stage, tile, (uint32_t)(it / S));
        const int np = ring.points(tile);
        const float* ztile = ring.stage_ptr(stage);
        // ---------------- phase 1: E-step for point threadIdx.x ----------------
        float rk[KP];                    // this point's responsibilities (ballots of the sparsity skip below)
        {
            const bool active = (int)threadIdx.x < np;
            float lp[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) lp[k] = 0.f;
            float lse = 0.f;
            int label = 0;
            if (active) {
                float x[D];
                load_row<D>(ztile, threadIdx.x, x);
                float best = -3.4e38f;
#pragma unroll 1
                for (int k = 0; k < K; ++k) {
                    float df[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) df[c] = x[c] - mu_s[k * D + c];
                    const float* uk_s = u_s + k * TRI;
                    float m = 0.f;
#pragma unroll
                    for (int b = 0; b < D; ++b) {
                        float y = 0.f;
#pragma unroll
                        for (int c = 0; c <= b; ++c) y = fmaf(df[c], uk_s[tri(b) + c], y);
                        m = fmaf(y, y, m);
                    }
                    const float v = fmaf(-0.5f, m, cst_s[k]);
                    r_s[threadIdx.x * KP + k] = v;             // park log-probabilities (dynamic k)
                    if (v > best) { best = v; label = k; }
                }
#pragma unroll
                for (int k = 0; k < KP; ++k) lp[k] = (k < K) ? r_s[threadIdx.x * KP + k] : -3.4e38f;
                float se = 0.f;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < K) se += expf(lp[k] - best);
                lse = best + logf(se);
                loglik += lse;
            }
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                float r = (active && k < K) ? expf(lp[k] - lse) : 0.f;
                if ((a.accumulate & 3) == SCC_GMM_HARD) r = (active && k == label) ? 1.f : 0.f;
                lp[k] = r;
            }
#pragma unroll
            for (int k = 0; k < KP; ++k) rk[k] = lp[k];
#pragma unroll
            for (int k = 0; k < KP; k += 4)
                *reinterpret_cast<float4*>(r_s + threadIdx.x * KP + k) = make_float4(lp[k], lp[k + 1], lp[k + 2], lp[k + 3]);
            if (active) {
                const size_t i = (size_t)tile * TILE + threadIdx.x;
                if (a.labels) a.labels[i] = label;
                if (a.resp) {
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if (k < K) a.resp[i * K + k] = lp[k];
                }
            }
        }
        // ---- responsibility-sparsity skip (see gmm_em_sparse_kernel): per-component ballots -> compact point lists
        if (a.accumulate & 3) {
            const bool active = (int)threadIdx.x < np;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const unsigned int mk = __ballot_sync(0xffffffffu, active && k < K && rk[k] >= thr && rk[k] > 0.f);
                if (lane == 0) mask_s[k * NWB + warp] = mk;
            }
        }
        __syncthreads();
        // ---------------- phase 2: 4x4 moment blocks over the component's significant points ----------------
        if (a.accumulate & 3) {
            if (threadIdx.x < KP * NWB) {
                const int k = threadIdx.x / NWB, w = threadIdx.x - k * NWB;
                int pfx = 0;
                for (int ww = 0; ww < w; ++ww) pfx += __popc(mask_s[k * NWB + ww]);
                pref_s[k * (NWB + 1) + w] = pfx;
                if (w == NWB - 1) pref_s[k * (NWB + 1) + NWB] = pfx + __popc(mask_s[k * NWB + w]);
            }
            __syncthreads();
            const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const unsigned int mk = mask_s[k * NWB + warp];
                if ((mk >> lane) & 1u)
                    list_s[k * TILE + pref_s[k * (NWB + 1) + warp] + __popc(mk & lt)] = (unsigned short)threadIdx.x;
            }
            __syncthreads();
#pragma unroll
            for (int m = 0; m < MAXU; ++m) {
                if (uk[m] >= 0) {
                    const int k = uk[m];
                    const int cnt = pref_s[k * (NWB + 1) + NWB];
                    for (int e = 0; e < cnt; ++e) {
                        const int t = list_s[k * TILE + e];
                        const float* row = ztile + t * L::LD;
                        const float r = r_s[t * KP + k];
                        const float4 xa = *reinterpret_cast<const float4*>(row + 4 * ua[m]);
                        const float4 xb = *reinterpret_cast<const float4*>(row + 4 * ub[m]);
                        const float da[4] = {xa.x - mua[m][0], xa.y - mua[m][1], xa.z - mua[m][2], xa.w - mua[m][3]};
                        const float db[4] = {xb.x - mub[m][0], xb.y - mub[m][1], xb.z - mub[m][2], xb.w - mub[m][3]};
#pragma unroll
                        for (int rr = 0; rr < 4; ++rr) {
                            const float w = r * da[rr];
                            s1[m][rr] += w;
#pragma unroll
                            for (int c = 0; c < 4; ++c) acc[m][4 * rr + c] = fmaf(w, db[c], acc[m][4 * rr + c]);
                        }
                        s0[m] += r;
                    }
                }
            }
            if ((it + 1) % FLUSH == 0) flush();
        }
        __syncthreads();
        ring.issue(stage, tile + S * G);
    }
    if (a.accumulate & 3) flush();
    {
        const float w = warp_sum(loglik);
        if (lane == 0) ll_s[warp] = (double)w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < NT / 32; ++w) v += ll_s[w];
        slot[0] = v;
    }
}

template <int D, int KP>
constexpr size_t gmm_block_smem() {
    using B = GmmBlock<D, KP>;
    return sizeof(float) * (B::S * B::TILE * RowLayout<D>::LD + B::TILE * KP + KP * D + ((KP * B::TRI + 3) & ~3) +
                            ((KP + 3) & ~3)) +
           sizeof(double) * (B::NT / 32) + sizeof(uint64_t) * B::S +
           sizeof(unsigned int) * KP * (B::NT / 32) + sizeof(int) * KP * (B::NT / 32 + 1) + sizeof(unsigned short) * KP * B::TILE;
}

template <int D, int KP>
static int launch_gmm_block(const GmmArgs& a, cudaStream_t st) {
    using B = GmmBlock<D, KP>;
    auto kern = gmm_em_block_kernel<D, KP>;
    constexpr size_t smem = gmm_block_smem<D, KP>();
    const int64_t tiles = (a.n + B::TILE - 1) / B::TILE;
    int64_t grid = persistent_grid(reinterpret_cast<const void*>(kern), B::NT, smem, 1);
    if (grid < 0) return (int)grid;
    if (grid > kMaxGmmGrid) grid = kMaxGmmGrid;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, B::NT, smem, st>>>(a);
    SCC_CUDA(cudaGetLastError());
    return gmm_after_stats(a, (int)grid, D, st);
}

template <int D, int KP>
static int launch_gmm(const GmmArgs& a, cudaStream_t st) {
    if constexpr (D <= 12) return launch_gmm_small<D, KP>(a, st);
    else return launch_gmm_block<D, KP>(a, st);
}

}  // namespace scc
