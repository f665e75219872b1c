// gmm_kernels.cuh — fused full-covariance EM pass for sm_100a (FP32 CUDA cores, packed FFMA2).
//
//   gmm_em_full_kernel<D,KP>   one read of z per EM iteration                     (d <= 12)
//       phase 1: one thread per point — Cholesky log-likelihoods of all K components
//                (sklearn _gaussian_mixture.py:490-553), log-sum-exp responsibilities
//                (_base.py:552-582), labels; r_ik parked in shared memory.
//                The mixture parameters live in CONSTANT memory: they reach the FFMA2s through
//                uniform registers (LDCU.128 on the uniform datapath), so the E-step issues no
//                shared-memory loads and no register splats
//                (SASS: FFMA2 R, R.F32, UR.F32x2.HI_LO, R.F32x2).
//       phase 2: one warp per component — the lanes sweep the tile's points and keep the
//                1 + d + d(d+1)/2 moments of "their" component in registers (as float2 pairs),
//                centred on the current mean (sklearn _gaussian_mixture.py:282-320,168-197).
//
// FP32-FMA-bound, not HBM-bound: ~K(d^2+4d) FMA per point against 4d bytes.
#pragma once

#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

__host__ __device__ constexpr int tri(int d) { return d * (d + 1) / 2; }

// Constant-memory layout of ONE component (floats), all pair-aligned:
//   [0, 2*DP2)            -mu_k as pairs (pad lane 0)
//   [2*DP2, 2*DP2+2*NU2)  U pairs: for bp in [0,DP2), c in [0, min(2bp+2, D)): {U[c][2bp], U[c][2bp+1]}
//                          (0 where c > b or b >= D)
//   [.., +2)              {cst_k, 0}
template <int D>
struct GmmConst {
    static constexpr int DP2 = (D + 1) / 2;
    __host__ __device__ static constexpr int rows(int bp) { return (2 * bp + 2 < D) ? 2 * bp + 2 : D; }
    __host__ __device__ static constexpr int upairs() { int n = 0; for (int bp = 0; bp < DP2; ++bp) n += rows(bp); return n; }
    __host__ __device__ static constexpr int uoff(int bp) { int n = 0; for (int b = 0; b < bp; ++b) n += rows(b); return n; }
    static constexpr int NU2 = upairs();
    static constexpr int kStride = 2 * DP2 + 2 * NU2 + 2;          // floats per component
};

constexpr int kGmmConstFloats = 16 * GmmConst<12>::kStride;         // largest instantiated shape
static __constant__ __align__(16) float c_gmm[kGmmConstFloats];

// ABI block (means[K*D] | U column-packed [K*TRI] | cst[K]) -> constant-memory layout (staged in global)
template <int D>
__global__ void gmm_repack_kernel(const float* __restrict__ params, int K, float* __restrict__ staged) {
    using C = GmmConst<D>;
    constexpr int TRI = tri(D);
    for (int i = threadIdx.x; i < K * C::kStride; i += blockDim.x) {
        const int k = i / C::kStride, o = i - k * C::kStride;
        float v = 0.f;
        if (o < 2 * C::DP2) {
            if (o < D) v = -params[k * D + o];
        } else if (o < 2 * C::DP2 + 2 * C::NU2) {
            const int e = o - 2 * C::DP2, pair = e >> 1, lane = e & 1;
            int bp = 0, base = 0;
            while (base + C::rows(bp) <= pair) { base += C::rows(bp); ++bp; }
            const int c = pair - base, b = 2 * bp + lane;
            if (b < D && c <= b) v = params[K * D + k * TRI + tri(b) + c];
        } else if (o == 2 * C::DP2 + 2 * C::NU2) {
            v = params[K * D + K * TRI + k];
        }
        staged[i] = v;
    }
}

struct GmmArgs {
    const float* z;
    int64_t n;
    int K;
    const float* params;       // ABI block (only used by the repack kernel)
    int32_t* labels;
    float* resp;
    const double* ctrl;
    int accumulate;            // 0 E-step only, 1 soft EM, 2 hard (one-hot) responsibilities
    double* stats;
    double* partials;
    unsigned int* counter;
};

// number of (aligned) b-pairs of row c of the symmetric second moment: bp in [c/2, DP2)
template <int D>
struct MomLayout {
    static constexpr int DP2 = (D + 1) / 2;
    __host__ __device__ static constexpr int off(int c) { int n = 0; for (int r = 0; r < c; ++r) n += DP2 - r / 2; return n; }
    static constexpr int NP = off(D);                       // pairs of S2
    static constexpr int NPAIR = 1 + DP2 + NP;              // {S0,0}, S1 pairs, S2 pairs
};

template <int D, int KP>
__global__ void __launch_bounds__(32 * KP, 1)
gmm_em_full_kernel(const GmmArgs a) {
    constexpr int NT = 32 * KP;
    constexpr int TILE = NT;
    constexpr int S = 3;
    constexpr int TRI = tri(D);
    constexpr int NM = 1 + D + TRI;                      // moments per component (ABI order)
    constexpr int FLUSH = 16;                            // tiles between float -> double flushes
    using Ring = ZRing<D, TILE, S, NT>;
    using C = GmmConst<D>;
    using M = MomLayout<D>;
    constexpr int DP2 = C::DP2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* r_s = ring_buf + S * Ring::kTileFloats;       // [KP][TILE]
    double* mom_s = reinterpret_cast<double*>(r_s + KP * TILE);   // [KP][NM]
    double* ll_s = mom_s + KP * NM;                      // [KP] per-warp log-likelihood
    double* cta_stats = ll_s + KP;                       // [1 + K*NM]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + 1 + KP * NM);

    if (a.ctrl && a.ctrl[5] != 0.0) return;              // frozen fit: converged or failed earlier

    const int K = a.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < KP * NM; i += NT) mom_s[i] = 0.0;

    Ring ring;
    ring.init(ring_buf, bars, a.z, a.n);
    __syncthreads();
    const int G = gridDim.x;
#pragma unroll
    for (int s = 0; s < S; ++s) ring.issue(s, blockIdx.x + s * G);
    __syncthreads();

    // phase-2 state of this warp's component (kc is warp-uniform; constant reads with a runtime
    // component index are plain LDCs, done once)
    const int kc = warp;
    float2 nmuk[DP2];
#pragma unroll
    for (int c = 0; c < DP2; ++c)
        nmuk[c] = make_float2(c_gmm[kc * C::kStride + 2 * c], c_gmm[kc * C::kStride + 2 * c + 1]);
    float2 mom2[M::NPAIR];
#pragma unroll
    for (int s = 0; s < M::NPAIR; ++s) mom2[s] = make_float2(0.f, 0.f);
    float loglik = 0.f;

    auto flush = [&]() {
        if (kc < K) {
            double* dst = mom_s + kc * NM;
            // S0
            { const float w = warp_sum(mom2[0].x); if (lane == 0) dst[0] += (double)w; }
            // S1
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float v = (c & 1) ? mom2[1 + (c >> 1)].y : mom2[1 + (c >> 1)].x;
                const float w = warp_sum(v);
                if (lane == 0) dst[1 + c] += (double)w;
            }
            // S2[c][b], c <= b, column-packed at tri(b) + c
#pragma unroll
            for (int c = 0; c < D; ++c) {
#pragma unroll
                for (int b = c; b < D; ++b) {
                    const int pr = 1 + DP2 + M::off(c) + (b >> 1) - (c >> 1);
                    const float v = (b & 1) ? mom2[pr].y : mom2[pr].x;
                    const float w = warp_sum(v);
                    if (lane == 0) dst[1 + D + tri(b) + c] += (double)w;
                }
            }
#pragma unroll
            for (int s = 0; s < M::NPAIR; ++s) mom2[s] = make_float2(0.f, 0.f);
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
                        const float* ck = c_gmm + k * C::kStride;           // compile-time offsets below
                        float2 df2[DP2];
#pragma unroll
                        for (int c = 0; c < DP2; ++c)
                            df2[c] = __fadd2_rn(x2[c], make_float2(ck[2 * c], ck[2 * c + 1]));
                        float2 m2 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int bp = 0; bp < DP2; ++bp) {
                            float2 y2 = make_float2(0.f, 0.f);
#pragma unroll
                            for (int c = 0; c < C::rows(bp); ++c) {
                                const float dc = (c & 1) ? df2[c >> 1].y : df2[c >> 1].x;
                                const int o = 2 * DP2 + 2 * (C::uoff(bp) + c);
                                y2 = __ffma2_rn(make_float2(dc, dc), make_float2(ck[o], ck[o + 1]), y2);
                            }
                            m2 = __ffma2_rn(y2, y2, m2);
                        }
                        lp[k] = fmaf(-0.5f, m2.x + m2.y, ck[2 * DP2 + 2 * C::NU2]);
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
                if (a.accumulate == SCC_GMM_HARD) r = (active && k == label) ? 1.f : 0.f;
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
        if (a.accumulate && kc < K) {
            for (int t = lane; t < np; t += 32) {
                const float r = r_s[kc * TILE + t];
                float xr[D];
                load_row<D>(ztile, t, xr);
                float2 df2[DP2], w2[DP2];
                const float2 r2 = splat2(r);
#pragma unroll
                for (int c = 0; c < DP2; ++c) {
                    df2[c] = __fadd2_rn(make_float2(xr[2 * c], (2 * c + 1 < D) ? xr[2 * c + 1] : 0.f), nmuk[c]);
                    w2[c] = __fmul2_rn(r2, df2[c]);                          // r (x - mu)
                    mom2[1 + c] = __fadd2_rn(mom2[1 + c], w2[c]);            // S1
                }
                mom2[0].x += r;                                              // S0
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const float wc = (c & 1) ? w2[c >> 1].y : w2[c >> 1].x;
#pragma unroll
                    for (int bp = c >> 1; bp < DP2; ++bp) {                  // S2[c][2bp..2bp+1]
                        const int pr = 1 + DP2 + M::off(c) + bp - (c >> 1);
                        mom2[pr] = __ffma2_rn(make_float2(wc, wc), df2[bp], mom2[pr]);
                    }
                }
            }
            if ((it + 1) % FLUSH == 0) flush();
        }
        __syncthreads();
        ring.issue(stage, tile + S * G);
    }
    if (a.accumulate) flush();
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
    return sizeof(float) * (S * TILE * RowLayout<D>::LD + KP * TILE) +
           sizeof(double) * (KP * NM + KP + 1 + KP * NM) + sizeof(uint64_t) * S;
}

template <int D, int KP>
static int launch_gmm_full(const GmmArgs& a, cudaStream_t st) {
    constexpr int NT = 32 * KP;
    using C = GmmConst<D>;
    static_assert(16 * C::kStride <= kGmmConstFloats, "constant buffer too small");
    auto kern = gmm_em_full_kernel<D, KP>;
    constexpr size_t smem = gmm_full_smem<D, KP>();
    const int NS = SCC_GMM_STAT_DOUBLES(a.K, D);
    const int64_t tiles = (a.n + NT - 1) / NT;
    int64_t grid = persistent_grid(reinterpret_cast<const void*>(kern), NT, smem, 2);
    if (grid < 0) return (int)grid;
    if (grid > kMaxGmmGrid) grid = kMaxGmmGrid;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    // stage the parameters in the constant-memory layout behind the partial slots, then copy them
    // into the constant bank (stream-ordered; one GMM pass per device at a time may be in flight)
    float* staged = reinterpret_cast<float*>(a.partials + (size_t)kMaxGmmGrid * NS);
    gmm_repack_kernel<D><<<1, 256, 0, st>>>(a.params, a.K, staged);
    SCC_CUDA(cudaGetLastError());
    SCC_CUDA(cudaMemcpyToSymbolAsync(c_gmm, staged, sizeof(float) * a.K * C::kStride, 0, cudaMemcpyDeviceToDevice, st));
    kern<<<(unsigned)grid, NT, smem, st>>>(a);
    SCC_CUDA(cudaGetLastError());
    reduce_partials_kernel<<<(NS + 255) / 256, 256, 0, st>>>(a.partials, NS, (int)grid, a.stats, a.ctrl);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

}  // namespace scc
