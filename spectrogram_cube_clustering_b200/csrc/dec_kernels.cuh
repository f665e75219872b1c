// dec_kernels.cuh — DEC clustering-layer kernel templates for sm_100a (CUDA cores, HBM-bound).
//
//   dec_assign_kernel   z -> q, labels, f_j = sum_i q_ij, label-change count   (one read of z)
//                       replaces Cluster/networks.py:279-288 + models.py:92,94,1098-1099,1320
//   dec_grad_reg_kernel / dec_grad_tiled_kernel
//                       z (+p | +f | +dL/dq) -> loss, dz, dmu                   (models.py:1124-1127 + autograd)
//       REG   variant: dmu accumulated in per-thread registers          (KP*D <= 160)
//       TILED variant: warp-level 4x4 register-blocked W^T Z over the staged tile (D % 4 == 0)
//
// One thread owns one latent point: its row sits in registers, centroids are broadcast from
// shared memory, q / coefficients never leave registers.  Template parameters:
//   D      latent dimension          KP     compile-time bound on the cluster count (4, 8, 16)
//   EXACT  K == KP (no per-cluster guards)      ALPHA1  alpha == 1 (no pow)
// Instantiated per dimension by dec_inst.cu (one translation unit per D, built in parallel).
#pragma once

#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

constexpr int kDecThreads = 256;
constexpr int kDecTile = 256;
constexpr int kBatchGridX = 296;       // grid.x bound of a batched (grid.y = restarts) Lloyd launch

// doubles of reduction scratch for an NV-long statistics vector: cta_reduce needs
// [num_warps][round_up(NV, 32)], grid_publish needs 2 * kDecThreads
__host__ __device__ constexpr int reduce_scratch(int nv) {
    return (kDecThreads / 32) * ((nv + 31) / 32 * 32) > 2 * kDecThreads ? (kDecThreads / 32) * ((nv + 31) / 32 * 32)
                                                                      : 2 * kDecThreads;
}

template <int D>
__host__ __device__ constexpr int dec_stages() { return RowLayout<D>::kDense ? 4 : (RowLayout<D>::kVec4 ? 3 : 2); }

// ---------------------------------------------------------------------------
// Packed FP32 (sm_100 FFMA2 / FADD2 / FMUL2): two lanes of a float2 per instruction.  On B200 a
// stream of 3-register scalar FFMAs issues at ~56 % of the FP32 peak (register-read bandwidth)
// while FFMA2 reaches ~88 % (tools/ubench_fp32.cu), and the issue-slot count halves.
// Rows are held as DP2 = ceil(D/2) float2 pairs; the pad lane of an odd D is kept at 0.
// ---------------------------------------------------------------------------
template <int D>
struct Pairs { static constexpr int N = (D + 1) / 2; };

template <int D>
__device__ __forceinline__ void pack_row(const float (&r)[D], float2 (&p)[Pairs<D>::N]) {
#pragma unroll
    for (int c = 0; c < Pairs<D>::N; ++c) p[c] = make_float2(r[2 * c], (2 * c + 1 < D) ? r[2 * c + 1] : 0.f);
}

// ---------------------------------------------------------------------------
// Student's-t kernel of one point against every centroid.  networks.py:279-288, models.py:92.
//   w_j = 1 + ||z - mu_j||^2 / alpha,  u_j = 1 / w_j,  t_j = u_j^((alpha+1)/2),  tsum = sum_j t_j
// so q_j = t_j / tsum.  Distances use the exact difference form (z_c - mu_jc)^2: no cancellation.
// nmu2_s holds the NEGATED centroids as float2 pairs [KP][DP2] (pad lane 0).
// LABEL: also the hard label = argmin distance (== argmax q, first index wins) and that distance.
// The arithmetic (operation order included) is the same as soft_assign_rows() of the assign kernel,
// so the gradient kernels recompute bit-identical q.  Reciprocals are single MUFU.RCP instructions (w >= 1, so no range fix-up is needed).
// ---------------------------------------------------------------------------
template <int D, int KP, bool EXACT, bool ALPHA1, bool LABEL>
__device__ __forceinline__ void student_t_row(const float2 (&z2)[Pairs<D>::N], const float2* __restrict__ nmu2_s,
                                              int K, float inv_alpha, float expo, float (&w)[KP], float (&u)[KP],
                                              float (&t)[KP], float& tsum, int& label, float& best) {
    constexpr int DP2 = Pairs<D>::N;
    float ts[2] = {0.f, 0.f};                 // two chains: short serial dependencies matter at 4 warps/scheduler
    best = 3.4e38f;
    label = 0;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        w[j] = 1.f; u[j] = 0.f; t[j] = 0.f;
        if (EXACT || j < K) {
            float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < DP2; ++c) {
                const float2 df = __fadd2_rn(z2[c], nmu2_s[j * DP2 + c]);
                acc2 = __ffma2_rn(df, df, acc2);
            }
            const float acc = acc2.x + acc2.y;
            if (LABEL) {
                if (acc < best) { best = acc; label = j; }
            }
            const float ww = ALPHA1 ? (1.f + acc) : fmaf(acc, inv_alpha, 1.f);
            const float uu = rcp_approx(ww);
            const float tt = ALPHA1 ? uu : ex2_approx(-expo * lg2_approx(ww));
            w[j] = ww; u[j] = uu; t[j] = tt; ts[j & 1] += tt;
        }
    }
    tsum = ts[0] + ts[1];
}

// negated centroids as pairs: nmu2_s[j][c] = -(mu[j][2c], mu[j][2c+1]); rows j >= K and pad lanes are 0
template <int D, int KP>
__device__ __forceinline__ void load_neg_centroid_pairs(const float* __restrict__ mu, int K, float2* nmu2_s) {
    constexpr int DP2 = Pairs<D>::N;
    float* flat = reinterpret_cast<float*>(nmu2_s);
    for (int i = threadIdx.x; i < KP * DP2 * 2; i += kDecThreads) {
        const int j = i / (2 * DP2), c = i - j * (2 * DP2);
        flat[i] = (j < K && c < D) ? -mu[j * D + c] : 0.f;
    }
}

template <int KP, bool EXACT>
__device__ __forceinline__ void store_krow(float* __restrict__ dst, int K, const float (&v)[KP]) {
    if (EXACT || (K & 3) == 0) {
#pragma unroll
        for (int j = 0; j < KP; j += 4)
            if (EXACT || j < K) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < KP; ++j)
            if (j < K) dst[j] = v[j];
    }
}
template <int KP, bool EXACT>
__device__ __forceinline__ void load_krow(const float* __restrict__ src, int K, float (&v)[KP]) {
    if (EXACT || (K & 3) == 0) {
#pragma unroll
        for (int j = 0; j < KP; j += 4) {
            if (EXACT || j < K) {
                const float4 x = ldg_stream4(reinterpret_cast<const float4*>(src + j));
                v[j] = x.x; v[j + 1] = x.y; v[j + 2] = x.z; v[j + 3] = x.w;
            } else { v[j] = 0.f; v[j + 1] = 0.f; v[j + 2] = 0.f; v[j + 3] = 0.f; }
        }
    } else {
#pragma unroll
        for (int j = 0; j < KP; ++j) v[j] = (j < K) ? ldg_stream(src + j) : 0.f;
    }
}

// Reduce NV per-thread floats across the CTA into cta_stats[0..NV) (float64).
// scratch: [num_warps][round_up(NV, 32)] doubles.  Deterministic (fixed tree / warp order).
template <int NV>
__host__ __device__ constexpr int reduce_pad() { return (NV + 31) / 32 * 32; }

template <int NV, int NT>
__device__ __forceinline__ void cta_reduce(const float (&v)[NV], double* scratch, double* cta_stats) {
    constexpr int NVP = reduce_pad<NV>();
    constexpr int M = NVP / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float x[NVP];
#pragma unroll
    for (int s = 0; s < NVP; ++s) x[s] = (s < NV) ? v[s] : 0.f;
    warp_reduce_scatter<NVP>(x);            // lane l now holds the warp totals of entries M*l .. M*l+M-1
#pragma unroll
    for (int r = 0; r < M; ++r) scratch[warp * NVP + M * lane + r] = (double)x[r];
    __syncthreads();
    for (int s = threadIdx.x; s < NV; s += NT) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) acc += scratch[w * NVP + s];
        cta_stats[s] = acc;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// P rows per thread sharing every centroid load: the centroid broadcasts (one LDS.64/128 per pair)
// are what saturates first at large K*d (LSU pipe: 1 wavefront/clk/SM against 4 FP32 warp-instr/clk),
// so register-blocking P = 2 points halves the shared-memory traffic per point.
// ---------------------------------------------------------------------------
template <int D, int KP, bool EXACT, bool ALPHA1, int P>
__device__ __forceinline__ void soft_assign_rows(const float2 (&z2)[P][Pairs<D>::N], const float2* __restrict__ nmu2_s,
                                                 int K, float inv_alpha, float expo,
                                                 float (&q)[P][KP], int (&label)[P]) {
    constexpr int DP2 = Pairs<D>::N;
    float tsum[P][2], best[P];
#pragma unroll
    for (int r = 0; r < P; ++r) { tsum[r][0] = 0.f; tsum[r][1] = 0.f; best[r] = 3.4e38f; label[r] = 0; }
#pragma unroll
    for (int j = 0; j < KP; ++j) {
#pragma unroll
        for (int r = 0; r < P; ++r) q[r][j] = 0.f;
        if (EXACT || j < K) {
            float2 acc2[P];
#pragma unroll
            for (int r = 0; r < P; ++r) acc2[r] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < DP2; ++c) {
                const float2 m = nmu2_s[j * DP2 + c];
#pragma unroll
                for (int r = 0; r < P; ++r) {
                    const float2 df = __fadd2_rn(z2[r][c], m);
                    acc2[r] = __ffma2_rn(df, df, acc2[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < P; ++r) {
                const float acc = acc2[r].x + acc2[r].y;
                if (acc < best[r]) { best[r] = acc; label[r] = j; }
                const float ww = ALPHA1 ? (1.f + acc) : fmaf(acc, inv_alpha, 1.f);
                const float t = ALPHA1 ? rcp_approx(ww) : ex2_approx(-expo * lg2_approx(ww));
                q[r][j] = t; tsum[r][j & 1] += t;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < P; ++r) {
        const float inv = rcp_approx(tsum[r][0] + tsum[r][1]);
#pragma unroll
        for (int j = 0; j < KP; ++j) q[r][j] *= inv;
    }
}

// points per thread of the assign pass (tile = 256 * P points)
template <int D, int KP>
__host__ __device__ constexpr int assign_ppt() { return (KP * D <= 96) ? 2 : 1; }
template <int D, int KP>
__host__ __device__ constexpr int assign_stages() {
    return RowLayout<D>::kDense ? (assign_ppt<D, KP>() == 2 ? 3 : 4) : (RowLayout<D>::kVec4 ? 3 : 2);
}

// ---------------------------------------------------------------------------
// dec_assign
// ---------------------------------------------------------------------------
template <int D, int KP, bool EXACT, bool ALPHA1>
__global__ void __launch_bounds__(kDecThreads)
dec_assign_kernel(const DecArgs a) {
    constexpr int P = assign_ppt<D, KP>();
    constexpr int TILE = kDecTile * P;
    constexpr int S = assign_stages<D, KP>();
    // CTA-level ring: a per-warp ring (WarpRing) was measured for this kernel too — its 6x more, 6x smaller
    // TMA copies lengthen the prologue by ~1.3 us and the short main loop gains nothing
    using Ring = ZRing<D, TILE, S, kDecThreads>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    constexpr int DP2 = Pairs<D>::N;
    float2* nmu2_s = reinterpret_cast<float2*>(ring_buf + S * Ring::kTileFloats);      // [KP][DP2] (-mu pairs)
    double* cta_stats = reinterpret_cast<double*>(nmu2_s + ((KP * DP2 + 1) & ~1));     // [KP+1]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + (KP + 1));                // [S]
    // the reduction scratch [reduce_scratch(KP+1)] reuses the ring once the main loop is over: keeping it
    // separate pushes the d = 32 kernel (3 x 36 KB of stages) over half an SM's shared memory -> 1 CTA/SM
    double* scratch = reinterpret_cast<double*>(ring_buf);
    static_assert(sizeof(float) * S * Ring::kTileFloats >= sizeof(double) * reduce_scratch(KP + 1), "ring too small");

    const int K = EXACT ? KP : a.K;
    Ring ring;
    ring.init(ring_buf, bars, a.z, a.n);
    __syncthreads();
    pdl_wait();                         // no global access before this point (see scc_common.cuh)
    SCC_TL(a.timeline, 0);
    const int G = gridDim.x;
#pragma unroll
    for (int s = 0; s < S; ++s) ring.issue(s, blockIdx.x + s * G);
    load_neg_centroid_pairs<D, KP>(a.mu, K, nmu2_s);
    __syncthreads();
    SCC_TL(a.timeline, 1);

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    const bool round5 = a.round5 != 0;
    float facc[KP + 1];
#pragma unroll
    for (int j = 0; j <= KP; ++j) facc[j] = 0.f;

    int stage = 0;
    uint32_t use = 0;
    for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G) {
        ring.wait(stage, tile, use);
        if (tile == (int)blockIdx.x) SCC_TL(a.timeline, 2);
        const int np = ring.points(tile);
        float2 z2[P][DP2];
        bool active[P];
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const int t = threadIdx.x + r * kDecThreads;
            active[r] = t < np;
            float zr[D];
#pragma unroll
            for (int c = 0; c < D; ++c) zr[c] = 0.f;
            if (active[r]) load_row<D>(ring.stage_ptr(stage), t, zr);
            pack_row<D>(zr, z2[r]);
        }
        __syncthreads();                 // every row of the stage is in registers: refill it
        ring.issue(stage, tile + S * G);
        if (active[0]) {
            float q[P][KP];
            int label[P];
            soft_assign_rows<D, KP, EXACT, ALPHA1, P>(z2, nmu2_s, K, inv_alpha, expo, q, label);
#pragma unroll
            for (int r = 0; r < P; ++r) {
                if (active[r]) {
                    const size_t i = (size_t)tile * TILE + threadIdx.x + r * kDecThreads;
                    if (round5) {
#pragma unroll
                        for (int j = 0; j < KP; ++j) q[r][j] = round_dec5(q[r][j]);
                    }
#pragma unroll
                    for (int j = 0; j < KP; ++j) facc[j] += q[r][j];
                    if (a.q) store_krow<KP, EXACT>(a.q + i * K, K, q[r]);
                    if (a.labels) a.labels[i] = label[r];
                    if (a.labels_prev) facc[KP] += (a.labels_prev[i] != label[r]) ? 1.f : 0.f;
                }
            }
        }
        if (++stage == S) { stage = 0; ++use; }
    }
    pdl_trigger();                      // successor may start its prologue under our reduction tail
    SCC_TL(a.timeline, 3);
    __syncthreads();                    // every warp is done with the ring: it becomes the reduction scratch
    cta_reduce<KP + 1, kDecThreads>(facc, scratch, cta_stats);
    SCC_TL(a.timeline, 4);
    if (!EXACT) {                       // stats layout is [K+1]: compact the KP-padded vector
        if (threadIdx.x == 0 && K < KP) cta_stats[K] = cta_stats[KP];
        __syncthreads();
    }
    const PeerCtx push{a.ex_push ? a.ex_windows : nullptr, a.ex_rank, a.ex_world, a.ex_max_len};
    grid_publish<kDecThreads>(cta_stats, K + 1, a.partials, a.counter, a.stats, scratch, &push);
    SCC_TL(a.timeline, 5);
}

// ---------------------------------------------------------------------------
// Per-point gradient coefficients c_ij with dz_i = cs sum_j c_ij (z_i - mu_j),
// dmu_j = -cs sum_i c_ij (z_i - mu_j).  The common factor cs is NOT applied here: it is folded into
// the -(mu - c0) table used for dz and into the final reduction of dmu (grad_fold_scale()).
//   MODE_KL/KLF  : c_ij = (p_ij - q_ij s_i) u_ij,  cs = scale (alpha+1)/alpha   (+ loss, in log2 units);
//                  KL streams p from memory, KLF rebuilds it from the column sums (and may write it out)
//   MODE_GENERIC : c_ij = q_ij (sum_j G_ij q_ij - G_ij) u_ij,  cs = (alpha+1)/alpha
//   MODE_KMEANS  : c_ij = [j == argmin_j ||z_i - mu_j||^2],  cs = 1  (Lloyd step: counts, centre shifts, inertia)
// Inputs are the Student's-t quantities of student_t_row(): q_j = t_j / tsum, 1/q_j = tsum w_j^expo.
// ---------------------------------------------------------------------------
// The [n, K] operand a gradient kernel streams besides z: the target p (MODE_KL, API mode) or the
// upstream gradient dL/dq (MODE_GENERIC).  The row is requested at the top of the iteration, before
// the wait on the z tile.  (Requesting it a whole iteration ahead was measured: the 8 extra live
// registers spill at the 128-register cap and the kernel got 12 % slower.)
template <int MODE>
__device__ __forceinline__ const float* krow_operand(const DecArgs& a) {
    return MODE == MODE_KL ? a.p : (MODE == MODE_GENERIC ? a.grad_q : nullptr);
}
template <int KP, bool EXACT>
__device__ __forceinline__ void prefetch_krow(const float* __restrict__ src, int64_t base, int np, int K,
                                              float (&row)[KP]) {
    if (src && (int)threadIdx.x < np) load_krow<KP, EXACT>(src + ((size_t)base + threadIdx.x) * K, K, row);
}

template <int MODE>
__host__ __device__ constexpr bool mode_is_kl() { return MODE == MODE_KL || MODE == MODE_KLF || MODE == MODE_STEP; }

template <int MODE>
__host__ __device__ __forceinline__ float grad_fold_scale(float scale, float alpha) {
    return MODE == MODE_KMEANS ? 1.f : (mode_is_kl<MODE>() ? scale : 1.f) * (alpha + 1.f) / alpha;
}

template <int KP, bool EXACT, bool ALPHA1, int MODE>
__device__ __forceinline__ void grad_coefficients(const DecArgs& a, size_t i, int K, const float* __restrict__ inv_f,
                                                  const float (&w)[KP], const float (&u)[KP], const float (&t)[KP],
                                                  float tsum, float expo, int label, float best,
                                                  const float (&pre)[KP],
                                                  float (&coef)[KP], float& loss, float& ssum) {
    if constexpr (MODE == MODE_KMEANS) {
        // Lloyd statistics: one-hot coefficient on the nearest centre, "loss" = inertia
#pragma unroll
        for (int j = 0; j < KP; ++j) coef[j] = (j == label) ? 1.f : 0.f;
        loss += best;
        if (a.labels) a.labels[i] = label;
        if (a.mindist) a.mindist[i] = best;
    } else if constexpr (mode_is_kl<MODE>()) {
        const float inv = rcp_approx(tsum);
        float p[KP];
        if constexpr (MODE == MODE_KL) {           // target row requested before the z tile wait (prefetch_krow)
#pragma unroll
            for (int j = 0; j < KP; ++j) p[j] = pre[j];
        } else {                                   // MODE_KLF / MODE_STEP: rebuild p from the column sums
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const float q = t[j] * inv;
                const float qq = a.round5 ? round_dec5(q) : q;
                p[j] = (EXACT || j < K) ? qq * qq * inv_f[j] : 0.f;
            }
            // row sum in the order dec_target_kernel uses (groups of 4, then a pairwise tree over the
            // groups; sequential when K is not 4, 8 or 16), so the rebuilt p is bit-identical to its output
            float wsum;
            if ((K & 3) == 0 && K != 12) {
                float g4[KP / 4];
#pragma unroll
                for (int b = 0; b < KP / 4; ++b) g4[b] = (p[4 * b] + p[4 * b + 1]) + (p[4 * b + 2] + p[4 * b + 3]);
                if constexpr (KP == 4) wsum = g4[0];
                else if constexpr (KP == 8) wsum = (K == 4) ? g4[0] : g4[0] + g4[1];
                else wsum = (K == 4) ? g4[0] : ((K == 8) ? g4[0] + g4[1] : (g4[0] + g4[1]) + (g4[2] + g4[3]));
            } else {
                wsum = 0.f;
#pragma unroll
                for (int j = 0; j < KP; ++j) wsum += (EXACT || j < K) ? p[j] : 0.f;
            }
            const float winv = 1.f / wsum;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                p[j] *= winv;
                if (a.round5) p[j] = round_dec5(p[j]);
            }
            if (a.p_out) store_krow<KP, EXACT>(a.p_out + i * K, K, p);      // materialise target_distribution(q)
        }
        // p_j / q_j = p_j w_j tsum for alpha == 1 (w_j = 1 + d_j is already in registers: no division); the log is
        // taken of the RATIO (near 1), not of its large factors separately — lg2.approx has a relative error, so
        // log2(p w) + log2(tsum) would lose the digits that cancel.  The 1e-37 keeps a zero target at
        // 0 * finite = 0 (torch KLDivLoss: xlogy); negative / NaN targets still give NaN.
        float s2[2] = {0.f, 0.f}, l2[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            if (EXACT || j < K) {
                s2[j & 1] += p[j];
                const float ratio = ALPHA1 ? fmaf(p[j] * w[j], tsum, 1e-37f)
                                           : fmaf(p[j], rcp_approx(fmaxf(t[j] * inv, 1e-37f)), 1e-37f);
                l2[j & 1] = fmaf(p[j], lg2_approx(ratio), l2[j & 1]);
            }
        }
        const float s = s2[0] + s2[1];
        const float nis = -(inv * s);
#pragma unroll
        for (int j = 0; j < KP; ++j) coef[j] = (EXACT || j < K) ? fmaf(t[j], nis, p[j]) * u[j] : 0.f;
        loss += l2[0] + l2[1];
        ssum += s;
    } else {
        const float inv = rcp_approx(tsum);
        float g[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) g[j] = pre[j];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < KP; ++j) dot = fmaf(g[j], t[j], dot);
        dot *= inv;
#pragma unroll
        for (int j = 0; j < KP; ++j) coef[j] = (EXACT || j < K) ? (t[j] * inv) * (dot - g[j]) * u[j] : 0.f;
    }
}

// dz_c = cs ((sum_j c_j) zc_c - sum_j c_j mc_jc)  with zc = z - c0, mc = mu - c0 (algebraic form of
// sum_j c_j (z_c - mu_jc): K*D/2 FFMA2 instead of K*D (FADD + FMA)).  nmc2_s = -cs (mu - c0) pairs,
// csum = cs sum_j c_j.  The pad lane of an odd D holds garbage and is never stored.
template <int D, int KP, bool EXACT>
__device__ __forceinline__ void dz_from_coefficients(const float2 (&zc2)[Pairs<D>::N], const float (&coef)[KP],
                                                     float csum, const float2* __restrict__ nmc2_s, int K,
                                                     float (&dzr)[D]) {
    constexpr int DP2 = Pairs<D>::N;
    float2 dz2[DP2];
    const float2 cs = splat2(csum);
#pragma unroll
    for (int c = 0; c < DP2; ++c) dz2[c] = __fmul2_rn(cs, zc2[c]);
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        if (EXACT || j < K) {
            const float2 cj = splat2(coef[j]);
#pragma unroll
            for (int c = 0; c < DP2; ++c) dz2[c] = __ffma2_rn(cj, nmc2_s[j * DP2 + c], dz2[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) dzr[c] = (c & 1) ? dz2[c >> 1].y : dz2[c >> 1].x;
}

// Coalesced copy of a staged [np, D] tile (row stride LD) to global rows.
template <int D, int NT>
__device__ __forceinline__ void copy_tile_out(const float* __restrict__ tile, float* __restrict__ dst, int np) {
    using L = RowLayout<D>;
    if constexpr (L::kVec4) {
        const int nvec = np * (D / 4);
        for (int v = threadIdx.x; v < nvec; v += NT) {
            const int row = v / (D / 4), c4 = v - row * (D / 4);
            reinterpret_cast<float4*>(dst)[v] = *reinterpret_cast<const float4*>(tile + row * L::LD + 4 * c4);
        }
    } else if constexpr (L::kDense) {
        const int nf = np * D;                       // dense tile: flat copy, 128-bit where aligned
        const int nvec = nf / 4;
        for (int v = threadIdx.x; v < nvec; v += NT)
            reinterpret_cast<float4*>(dst)[v] = reinterpret_cast<const float4*>(tile)[v];
        for (int f = nvec * 4 + threadIdx.x; f < nf; f += NT) dst[f] = tile[f];
    } else {
        const int nf = np * D;
        for (int f = threadIdx.x; f < nf; f += NT) {
            const int row = f / D, c = f - row * D;
            dst[f] = tile[row * L::LD + c];
        }
    }
}

// Shared prologue of the gradient kernels: -mu pairs, -cs (mu - c0) pairs, c0, (mu - c0), 1/f.
template <int D, int KP>
__device__ __forceinline__ void load_grad_constants(const DecArgs& a, int K, float cs, float2* nmu2_s,
                                                    float2* nmc2_s, float* mc_s, float* c0_s, float* inv_f) {
    constexpr int DP2 = Pairs<D>::N;
    if (threadIdx.x < D) {
        float m = 0.f;
        for (int j = 0; j < K; ++j) m += a.mu[j * D + threadIdx.x];
        c0_s[threadIdx.x] = m / (float)K;
    }
    if (a.ex_pull_f && a.ex_windows) {          // column sums arrive through the fused exchange
        __shared__ double f_pull[SCC_MAX_K + 1];
        const PeerCtx ex{a.ex_windows, a.ex_rank, a.ex_world, a.ex_max_len};
        peer_pull(ex, f_pull, K + 1);
        if (threadIdx.x < KP) inv_f[threadIdx.x] = ((int)threadIdx.x < K) ? (float)(1.0 / f_pull[threadIdx.x]) : 0.f;
    } else if (threadIdx.x < KP) {
        inv_f[threadIdx.x] = (a.f_cols && (int)threadIdx.x < K) ? (float)(1.0 / a.f_cols[threadIdx.x]) : 0.f;
    }
    __syncthreads();
    float* nmu = reinterpret_cast<float*>(nmu2_s);
    float* nmc = reinterpret_cast<float*>(nmc2_s);
    for (int i = threadIdx.x; i < KP * DP2 * 2; i += kDecThreads) {
        const int j = i / (2 * DP2), c = i - j * (2 * DP2);
        const bool ok = (j < K && c < D);
        const float m = ok ? a.mu[j * D + c] : 0.f;
        nmu[i] = -m;
        nmc[i] = ok ? -cs * (m - c0_s[c]) : 0.f;
    }
    for (int i = threadIdx.x; i < KP * D; i += kDecThreads) mc_s[i] = (i < K * D) ? a.mu[i] - c0_s[i % D] : 0.f;
}

// ---------------------------------------------------------------------------
// Batched Lloyd step (MODE_KMEANS only): gridDim.y independent restarts scan the SAME z against their own
// centres (KMeans(n_init=100), models.py:386-394).  blockIdx.y selects the restart: centres, statistics,
// partial slots, ticket counter and the optional per-point outputs are offset; restarts whose `done` flag is
// set (converged earlier) return at once.  Everything else in the kernel is unchanged: it only ever uses
// blockIdx.x / gridDim.x.  Returns false when this CTA has nothing to do.
// ---------------------------------------------------------------------------
template <int MODE, int D>
__device__ __forceinline__ bool batch_view(DecArgs& a, int K) {
    if constexpr (MODE == MODE_KMEANS) {
        if (a.batch > 0) {
            const int r = blockIdx.y;
            if (a.batch_done && a.batch_done[r]) return false;
            const int S = K * D + 2 + K;
            a.mu += (size_t)r * K * D;
            a.stats += (size_t)r * S;
            a.partials += (size_t)r * gridDim.x * ((S + 1) & ~1);
            a.counter += 2 * r;
            if (a.labels) a.labels += (size_t)r * a.n;
            if (a.mindist) a.mindist += (size_t)r * a.n;
        }
    }
    return true;
}

// ---------------------------------------------------------------------------
// dec_grad, REG variant.  Per-thread accumulators: loss, sum s, W_j = sum_i c_ij,
// B_jc = sum_i c_ij (z_ic - c0_c);  dmu_jc = -cs (B_jc - W_j (mu_jc - c0_c)).
// For odd D the pad lane of the last float2 pair of the centred point is the constant 1, so W_j
// accumulates in the pad lane of B_j for free (no separate registers / adds).
// stats out: [loss, sum_i s_i, dmu[K*D]]
// ---------------------------------------------------------------------------
template <int D, int KP, bool EXACT, bool ALPHA1, int MODE>
__global__ void __launch_bounds__(kDecThreads, (2 + KP + 2 * KP * Pairs<D>::N) <= 100 ? 2 : 1)
dec_grad_reg_kernel(const DecArgs a_in) {
    constexpr int S = dec_stages<D>();
    constexpr int DP2 = Pairs<D>::N;
    constexpr bool kPadW = (D & 1) != 0;
    using Ring = WarpRing<D, kDecTile, S, kDecThreads>;
    using L = RowLayout<D>;
    constexpr int NV = 2 + KP + KP * D;                                      // [loss, sum s, W[KP], B[KP*D]]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    // dz staging, per warp: dense row layouts ship the warp's 32 rows with a TMA bulk store from two
    // alternating buffers; padded layouts keep one buffer and a coalesced copy by the warp's lanes
    constexpr bool kBulkOut = L::kDense;
    constexpr int NOUT = kBulkOut ? 2 : 1;
    float* out_tile = ring_buf + S * Ring::kTileFloats;                      // [NOUT][TILE*LD]
    float2* nmu2_s = reinterpret_cast<float2*>(out_tile + NOUT * Ring::kTileFloats);  // [KP][DP2]
    float2* nmc2_s = nmu2_s + ((KP * DP2 + 1) & ~1);                         // [KP][DP2]
    float* mc_s = reinterpret_cast<float*>(nmc2_s + ((KP * DP2 + 1) & ~1));  // [KP*D]
    float* c0_s = mc_s + ((KP * D + 3) & ~3);                                // [D]
    float* inv_f = c0_s + ((D + 3) & ~3);                                    // [KP]
    float2* nc0_s = reinterpret_cast<float2*>(inv_f + ((KP + 3) & ~3));      // [DP2] -c0 pairs (pad lane: +1)
    double* scratch = reinterpret_cast<double*>(nc0_s + ((DP2 + 1) & ~1));   // [reduce_scratch(NV)]
    double* cta_stats = scratch + reduce_scratch(NV);                        // [NV]  (>= K*D + 2 + K)
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + NV);            // [NW][S]

    const int K = EXACT ? KP : a_in.K;
    Ring ring;
    ring.init(ring_buf, bars, a_in.z, a_in.n);
    pdl_wait();                         // no global access before this point (see scc_common.cuh)
    DecArgs a_view = a_in;
    if (!batch_view<MODE, D>(a_view, K)) return;
    const DecArgs& a = (MODE == MODE_KMEANS) ? a_view : a_in;
    const float cs = grad_fold_scale<MODE>(a.scale, a.alpha);
    SCC_TL(a.timeline, 0);
    const int G = gridDim.x;
#pragma unroll
    for (int s = 0; s < S; ++s) ring.issue(s, blockIdx.x + s * G);
    load_grad_constants<D, KP>(a, K, cs, nmu2_s, nmc2_s, mc_s, c0_s, inv_f);
    __syncthreads();
    SCC_TL(a.timeline, 1);

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    const bool want_dz = a.dz != nullptr;
    float sm[2 + (kPadW ? 0 : KP)];       // loss, sum s, (W_j when there is no pad lane)
#pragma unroll
    for (int s = 0; s < 2 + (kPadW ? 0 : KP); ++s) sm[s] = 0.f;
    float2 B2[KP * DP2];                  // B_jc = sum_i c_ij (z_ic - c0_c), as pairs
#pragma unroll
    for (int s = 0; s < KP * DP2; ++s) B2[s] = make_float2(0.f, 0.f);
    // -c0 as pairs, read back from shared memory once per tile (keeping them in registers costs
    // 2*DP2 live registers the accumulators need)
    if (threadIdx.x < DP2)
        nc0_s[threadIdx.x] = make_float2(-c0_s[2 * threadIdx.x],
                                         (2 * threadIdx.x + 1 < D) ? -c0_s[2 * threadIdx.x + 1] : 1.f);
    __syncthreads();

    const float* krows = krow_operand<MODE>(a);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int stage = 0, ob = 0;

    if constexpr (MODE == MODE_STEP) {
        // ---- pass 1 of the one-kernel DEC step: the assign pass (same arithmetic as dec_assign_kernel) over
        // this CTA's tiles, then a grid-wide barrier + all-reduce of f, all CTAs co-resident (cooperative launch)
        float facc[KP + 1];
#pragma unroll
        for (int j = 0; j <= KP; ++j) facc[j] = 0.f;
        for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G) {
            const int np = ring.points(tile);
            const bool active = (int)threadIdx.x < np;
            ring.wait(stage, tile);
            float zr[D];
            if (active) load_row<D>(ring.stage_ptr(stage), threadIdx.x, zr);
            __syncwarp();
            ring.issue(stage, tile + S * G);
            if (active) {
                const size_t i = (size_t)tile * kDecTile + threadIdx.x;
                float w[KP], u[KP], t[KP];
                float2 z2[DP2];
                pack_row<D>(zr, z2);
                int label;
                float best, tsum;
                student_t_row<D, KP, EXACT, ALPHA1, true>(z2, nmu2_s, K, inv_alpha, expo, w, u, t, tsum, label, best);
                const float inv = rcp_approx(tsum);
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    float q = t[j] * inv;
                    if (a.round5) q = round_dec5(q);
                    t[j] = q;
                    facc[j] += q;
                }
                if (a.q) store_krow<KP, EXACT>(a.q + i * K, K, t);
                if (a.labels) a.labels[i] = label;
                if (a.labels_prev) facc[KP] += (a.labels_prev[i] != label) ? 1.f : 0.f;
            }
            if (++stage == S) stage = 0;
        }
        SCC_TL(a.timeline, 6);                                     // pass 1 main loop done
        // pass 2's first tiles are requested before the grid barrier: they land while the CTAs wait
#pragma unroll
        for (int s = 0; s < S; ++s) ring.issue((stage + s) % S, blockIdx.x + s * G);
        __syncthreads();
        double* f_s = cta_stats;                                   // [K+1] (cta_stats is free until the tail)
        double* mine_s = cta_stats + ((KP + 2) & ~1);              // [KP+1] this CTA's sums
        cta_reduce<KP + 1, kDecThreads>(facc, scratch, mine_s);
        if (!EXACT) {
            if (threadIdx.x == 0 && K < KP) mine_s[K] = mine_s[KP];
            __syncthreads();
        }
        const int sp_tail = (K * D + 2 + 1) & ~1;                  // pass-1 slots live behind the tail's slots
        const PeerCtx ex1{a.ex_windows, a.ex_rank, a.ex_world, a.ex_max_len};
        grid_barrier_sum<kDecThreads>(mine_s, K + 1, a.partials + (size_t)gridDim.x * sp_tail, a.counter + 1, f_s,
                                      scratch, &ex1);
        if (threadIdx.x < KP) inv_f[threadIdx.x] = ((int)threadIdx.x < K) ? (float)(1.0 / f_s[threadIdx.x]) : 0.f;
        if (blockIdx.x == 0 && (int)threadIdx.x <= K && a.f_out) a.f_out[threadIdx.x] = f_s[threadIdx.x];
        __syncthreads();
        SCC_TL(a.timeline, 7);                                     // grid barrier + f all-reduce done
    }

    for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G) {       // no CTA-wide barrier in this loop
        const int np = ring.points(tile);
        const bool active = (int)threadIdx.x < np;
        // the [n, K] operand row (target p / upstream dL/dq) is requested before the wait on the z tile
        float kcur[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) kcur[j] = 0.f;
        prefetch_krow<KP, EXACT>(krows, (int64_t)tile * kDecTile, np, K, kcur);
        ring.wait(stage, tile);
        if (tile == (int)blockIdx.x) SCC_TL(a.timeline, 2);
        float zr[D];
        if (active) load_row<D>(ring.stage_ptr(stage), threadIdx.x, zr);
        __syncwarp();                    // the warp's rows are in registers (and its previous dz rows copied out)
        ring.issue(stage, tile + S * G);
        float* out_cur = out_tile + ob * Ring::kTileFloats;
        if (active) {
            const size_t i = (size_t)tile * kDecTile + threadIdx.x;
            float w[KP], u[KP], t[KP], coef[KP];
            float2 z2[DP2];
            pack_row<D>(zr, z2);
            int label;
            float best, tsum;
            student_t_row<D, KP, EXACT, ALPHA1, MODE == MODE_KMEANS>(z2, nmu2_s, K, inv_alpha, expo, w, u, t, tsum,
                                                                     label, best);
            grad_coefficients<KP, EXACT, ALPHA1, MODE>(a, i, K, inv_f, w, u, t, tsum, expo, label, best, kcur, coef,
                                                       sm[0], sm[1]);
            if constexpr (!kPadW) {
#pragma unroll
                for (int j = 0; j < KP; ++j) sm[2 + j] += coef[j];
            }
#pragma unroll
            for (int c = 0; c < DP2; ++c) z2[c] = __fadd2_rn(z2[c], nc0_s[c]);     // centred point (pad lane: 1)
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                if (EXACT || j < K) {
                    const float2 cj = splat2(coef[j]);
#pragma unroll
                    for (int c = 0; c < DP2; ++c) B2[j * DP2 + c] = __ffma2_rn(cj, z2[c], B2[j * DP2 + c]);
                }
            }
            if (want_dz) {
                float cs2[2] = {0.f, 0.f};
#pragma unroll
                for (int j = 0; j < KP; ++j) cs2[j & 1] += coef[j];
                float dzr[D];
                dz_from_coefficients<D, KP, EXACT>(z2, coef, cs * (cs2[0] + cs2[1]), nmc2_s, K, dzr);
                store_row<D>(out_cur, threadIdx.x, dzr);
            }
        }
        if (want_dz) {
            const int nv = (np == kDecTile) ? 32 : ring.slice_rows(np, 0);
            const float* src_w = out_cur + 32 * warp * L::LD;
            float* dst_w = a.dz + ((size_t)tile * kDecTile + 32 * warp) * D;
            if (kBulkOut && (np == kDecTile || Ring::tma_ok(nv))) {
                // the warp's rows -> async proxy -> one bulk store by lane 0.  wait_group.read 1 leaves this
                // store in flight but guarantees the previous one has finished reading the OTHER buffer,
                // which the warp overwrites after the next iteration's __syncwarp().
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    bulk_s2g(dst_w, src_w, (uint32_t)nv * D * sizeof(float));
                    bulk_commit();
                    bulk_wait_read<1>();
                }
                ob ^= 1;
            } else if (nv > 0) {
                __syncwarp();
                warp_copy_rows_out<D>(src_w, dst_w, nv);
            }
        }
        if (++stage == S) stage = 0;
    }
    if (kBulkOut && lane == 0) bulk_wait0();             // this warp's dz stores are complete
    pdl_trigger();                      // successor may start its prologue under our reduction tail
    SCC_TL(a.timeline, 3);
    if (mode_is_kl<MODE>()) sm[0] *= a.scale * 0.693147180559945f;     // loss = scale * ln2 * sum p log2(p/q)
    float acc[NV];
    acc[0] = sm[0]; acc[1] = sm[1];
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        if constexpr (kPadW) acc[2 + j] = B2[j * DP2 + DP2 - 1].y;
        else acc[2 + j] = sm[2 + (kPadW ? 0 : j)];
    }
#pragma unroll
    for (int j = 0; j < KP; ++j)
#pragma unroll
        for (int c = 0; c < D; ++c)
            acc[2 + KP + j * D + c] = (c & 1) ? B2[j * DP2 + (c >> 1)].y : B2[j * DP2 + (c >> 1)].x;
    __syncthreads();
    cta_reduce<NV, kDecThreads>(acc, scratch, cta_stats);
    SCC_TL(a.timeline, 4);
    // dmu_jc = -cs (B_jc - W_j (mu_jc - c0_c)), compacted to [loss, sum s, dmu[K*D]]
    // (MODE_KMEANS: [inertia, 0, sum_{i in j} (z_i - mu_j) [K*D], counts[K]])
    double dmu = 0.0, wj = 0.0;
    const int o = threadIdx.x;
    if (o < K * D) dmu = -(double)cs * (cta_stats[2 + KP + o] - cta_stats[2 + o / D] * (double)mc_s[o]);
    if (o < K) wj = cta_stats[2 + o];
    __syncthreads();
    if (o < K * D) cta_stats[2 + o] = (MODE == MODE_KMEANS) ? -dmu : dmu;
    if (MODE == MODE_KMEANS && o < K) cta_stats[2 + K * D + o] = wj;
    __syncthreads();
    const PeerCtx push{a.ex_push ? a.ex_windows : nullptr, a.ex_rank, a.ex_world, a.ex_max_len};
    const bool last = grid_publish<kDecThreads, 25>(cta_stats, K * D + 2 + (MODE == MODE_KMEANS ? K : 0), a.partials,
                                                    a.counter, a.stats, scratch, &push, a.ex_push);
    if (MODE == MODE_STEP && last && threadIdx.x == 0) a.counter[1] = 0u;    // every CTA is past the pass-1 barrier
    SCC_TL(a.timeline, 5);
}

// ---------------------------------------------------------------------------
// dec_grad, TILED variant (D % 4 == 0, KP % 4 == 0): phase 1 = thread per point
// (coefficients + dz), phase 2 = each warp accumulates W^T (Z - c0) over its
// share of the tile with 4x4 register blocks (one LDS.128 of W and one of Z per
// 16 FMAs).  c0 = mean centroid, removed to keep the sums well conditioned:
//   dmu_jc = -( A_jc - Wsum_j (mu_jc - c0_c) ),  A = sum_i c_ij (z_ic - c0_c).
// ---------------------------------------------------------------------------
template <int D, int KP, bool EXACT, bool ALPHA1, int MODE>
__global__ void __launch_bounds__(kDecThreads)
dec_grad_tiled_kernel(const DecArgs a_in) {
    static_assert(D % 4 == 0 && KP % 4 == 0, "tiled variant needs 4-aligned shapes");
    constexpr int S = 2;
    using Ring = ZRing<D, kDecTile, S, kDecThreads>;
    using L = RowLayout<D>;
    constexpr int NB = (KP / 4) * (D / 4);            // 4x4 output blocks
    constexpr int G2 = (32 / NB) > 0 ? (32 / NB) : 1; // point groups per warp
    constexpr int NW = kDecThreads / 32;
    constexpr int NSM = KP + 2;                       // loss, sum s, Wsum_j
    constexpr int NS = KP * D + 2;
    static_assert(NB <= 32, "too many output blocks for one warp");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* out_tile = ring_buf + S * Ring::kTileFloats;      // [TILE*LD] dz staging
    float* w_tile = out_tile + Ring::kTileFloats;            // [TILE*KP] coefficients
    constexpr int DP2 = Pairs<D>::N;
    float2* nmu2_s = reinterpret_cast<float2*>(w_tile + kDecTile * KP);      // [KP][DP2]
    float2* nmc2_s = nmu2_s + KP * DP2;                      // [KP][DP2]
    float* mc_s = reinterpret_cast<float*>(nmc2_s + KP * DP2);               // [KP*D]
    float* c0_s = mc_s + KP * D;                             // [D]
    float* inv_f = c0_s + D;                                 // [KP]
    double* scratch = reinterpret_cast<double*>(inv_f + KP); // [reduce_scratch(NSM)]
    double* small_s = scratch + reduce_scratch(NSM);         // [NSM]
    double* cta_stats = small_s + NSM;                       // [NS + KP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + NS + KP);

    const int K = EXACT ? KP : a_in.K;
    Ring ring;
    ring.init(ring_buf, bars, a_in.z, a_in.n);
    __syncthreads();
    pdl_wait();                         // no global access before this point (see scc_common.cuh)
    DecArgs a_view = a_in;
    if (!batch_view<MODE, D>(a_view, K)) return;
    const DecArgs& a = (MODE == MODE_KMEANS) ? a_view : a_in;
    const float cs = grad_fold_scale<MODE>(a.scale, a.alpha);
    const int G = gridDim.x;
#pragma unroll
    for (int s = 0; s < S; ++s) ring.issue(s, blockIdx.x + s * G);
    load_grad_constants<D, KP>(a, K, cs, nmu2_s, nmc2_s, mc_s, c0_s, inv_f);
    __syncthreads();

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    const bool want_dz = a.dz != nullptr;
    float2 nc0[DP2];
#pragma unroll
    for (int c = 0; c < DP2; ++c) nc0[c] = make_float2(-c0_s[2 * c], -c0_s[2 * c + 1]);
    float small[NSM];
#pragma unroll
    for (int s = 0; s < NSM; ++s) small[s] = 0.f;
    float blk[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) blk[s] = 0.f;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / NB, lb = lane - grp * NB;  // lanes >= G2*NB idle in phase 2
    const int jb = lb / (D / 4), cb = lb - jb * (D / 4);
    const bool p2_active = grp < G2;

    const float* krows = krow_operand<MODE>(a);
    int stage = 0;
    uint32_t use = 0;
    for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G) {
        const int np = ring.points(tile);
        const int64_t base = (int64_t)tile * kDecTile;
        float kcur[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) kcur[j] = 0.f;
        prefetch_krow<KP, EXACT>(krows, base, np, K, kcur);
        ring.wait(stage, tile, use);
        const bool active = (int)threadIdx.x < np;
        float* ztile = ring.stage_ptr(stage);
        float coef[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) coef[j] = 0.f;
        if (active) {
            float zr[D];
            load_row<D>(ztile, threadIdx.x, zr);
            const size_t i = (size_t)base + threadIdx.x;
            float w[KP], u[KP], t[KP];
            float2 z2[DP2];
            pack_row<D>(zr, z2);
            int label;
            float best, tsum;
            student_t_row<D, KP, EXACT, ALPHA1, MODE == MODE_KMEANS>(z2, nmu2_s, K, inv_alpha, expo, w, u, t, tsum,
                                                                     label, best);
            grad_coefficients<KP, EXACT, ALPHA1, MODE>(a, i, K, inv_f, w, u, t, tsum, expo, label, best, kcur, coef,
                                                       small[0], small[1]);
            float csum = 0.f;
#pragma unroll
            for (int j = 0; j < KP; ++j) { small[2 + j] += coef[j]; csum += coef[j]; }
#pragma unroll
            for (int c = 0; c < DP2; ++c) z2[c] = __fadd2_rn(z2[c], nc0[c]);
#pragma unroll
            for (int c = 0; c < D; ++c) zr[c] = (c & 1) ? z2[c >> 1].y : z2[c >> 1].x;
            store_row<D>(ztile, threadIdx.x, zr);     // own row, centred, for phase 2
            if (want_dz) {
                float dzr[D];
                dz_from_coefficients<D, KP, EXACT>(z2, coef, cs * csum, nmc2_s, K, dzr);
                store_row<D>(out_tile, threadIdx.x, dzr);
            }
        }
#pragma unroll
        for (int j = 0; j < KP; j += 4)
            *reinterpret_cast<float4*>(w_tile + threadIdx.x * KP + j) =
                make_float4(coef[j], coef[j + 1], coef[j + 2], coef[j + 3]);
        __syncthreads();
        if (want_dz) copy_tile_out<D, kDecThreads>(out_tile, a.dz + (size_t)base * D, np);
        if (p2_active) {
            for (int r = warp * G2 + grp; r < np; r += NW * G2) {
                const float4 w = *reinterpret_cast<const float4*>(w_tile + r * KP + 4 * jb);
                const float4 x = *reinterpret_cast<const float4*>(ztile + r * L::LD + 4 * cb);
                // scalar FMAs on purpose: the packed form (8 FFMA2 with a broadcast coefficient) was measured
                // 7 % SLOWER for d = 32, K = 16 — this phase is bound by its two LDS.128 per 16 FMAs
                blk[0] = fmaf(w.x, x.x, blk[0]);  blk[1] = fmaf(w.x, x.y, blk[1]);
                blk[2] = fmaf(w.x, x.z, blk[2]);  blk[3] = fmaf(w.x, x.w, blk[3]);
                blk[4] = fmaf(w.y, x.x, blk[4]);  blk[5] = fmaf(w.y, x.y, blk[5]);
                blk[6] = fmaf(w.y, x.z, blk[6]);  blk[7] = fmaf(w.y, x.w, blk[7]);
                blk[8] = fmaf(w.z, x.x, blk[8]);  blk[9] = fmaf(w.z, x.y, blk[9]);
                blk[10] = fmaf(w.z, x.z, blk[10]); blk[11] = fmaf(w.z, x.w, blk[11]);
                blk[12] = fmaf(w.w, x.x, blk[12]); blk[13] = fmaf(w.w, x.y, blk[13]);
                blk[14] = fmaf(w.w, x.z, blk[14]); blk[15] = fmaf(w.w, x.w, blk[15]);
            }
        }
        __syncthreads();                 // phase 2 done: stage, w_tile and out_tile reusable
        ring.issue(stage, tile + S * G);
        if (++stage == S) { stage = 0; ++use; }
    }
    // ---- CTA reduction ----
    pdl_trigger();
    if (mode_is_kl<MODE>()) small[0] *= a.scale * 0.693147180559945f;      // loss = scale * ln2 * sum p log2(p/q)
    cta_reduce<NSM, kDecThreads>(small, scratch, small_s);
    // per-(warp, group) 4x4 partials -> shared (the ring buffer is free now), fixed-order sum
    double* part = reinterpret_cast<double*>(ring_buf);                      // [NW*G2][KP*D]
    if (p2_active) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                part[(size_t)(warp * G2 + grp) * (KP * D) + (4 * jb + r) * D + 4 * cb + c] = (double)blk[4 * r + c];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < K * D; o += kDecThreads) {
        double accd = 0.0;
#pragma unroll
        for (int g = 0; g < NW * G2; ++g) accd += part[(size_t)g * (KP * D) + o];
        const double v = accd - small_s[2 + o / D] * (double)mc_s[o];
        cta_stats[2 + o] = (MODE == MODE_KMEANS) ? v : -(double)cs * v;
    }
    if (MODE == MODE_KMEANS && (int)threadIdx.x < K) cta_stats[2 + K * D + threadIdx.x] = small_s[2 + threadIdx.x];
    if (threadIdx.x == 0) { cta_stats[0] = small_s[0]; cta_stats[1] = small_s[1]; }
    __syncthreads();
    const PeerCtx push{a.ex_push ? a.ex_windows : nullptr, a.ex_rank, a.ex_world, a.ex_max_len};
    grid_publish<kDecThreads>(cta_stats, K * D + 2 + (MODE == MODE_KMEANS ? K : 0), a.partials, a.counter, a.stats,
                              scratch, &push, a.ex_push);
}

// ---------------------------------------------------------------------------
// Shared-memory footprints and launchers (per instantiation)
// ---------------------------------------------------------------------------
template <int D, int KP>
constexpr size_t assign_smem() {
    constexpr int S = assign_stages<D, KP>();
    return sizeof(float) * (S * kDecTile * assign_ppt<D, KP>() * RowLayout<D>::LD + 2 * ((KP * Pairs<D>::N + 1) & ~1)) +
           sizeof(double) * (KP + 1) + sizeof(uint64_t) * S;
}
template <int D, int KP>
constexpr size_t grad_reg_smem() {
    constexpr int S = dec_stages<D>();
    constexpr int NV = 2 + KP + KP * D;
    constexpr int SCR = reduce_scratch(NV);
    return sizeof(float) * ((S + (RowLayout<D>::kDense ? 2 : 1)) * kDecTile * RowLayout<D>::LD +
                            4 * ((KP * Pairs<D>::N + 1) & ~1) +
                            ((KP * D + 3) & ~3) + ((D + 3) & ~3) + ((KP + 3) & ~3) + 2 * ((Pairs<D>::N + 1) & ~1)) +
           sizeof(double) * (SCR + NV) + sizeof(uint64_t) * S * (kDecThreads / 32);
}
template <int D, int KP>
constexpr size_t grad_tiled_smem() {
    constexpr int S = 2;
    constexpr int NB = (KP / 4) * (D / 4);
    constexpr int G2 = (32 / NB) > 0 ? (32 / NB) : 1;
    constexpr int NW = kDecThreads / 32;
    constexpr int NSM = KP + 2;
    constexpr int scr = reduce_scratch(NSM);
    size_t bytes = sizeof(float) * ((S + 1) * kDecTile * RowLayout<D>::LD + kDecTile * KP + 4 * KP * Pairs<D>::N +
                                    KP * D + D + KP) +
                   sizeof(double) * (scr + NSM + KP * D + 2 + KP) + sizeof(uint64_t) * S;
    // the ring buffer is reused for the [NW*G2][KP*D] float64 partials at the end
    const size_t part = sizeof(double) * NW * G2 * KP * D;
    const size_t ring = sizeof(float) * S * kDecTile * RowLayout<D>::LD;
    if (part > ring) bytes += part - ring;
    return bytes;
}

template <typename Kern>
static int launch_dec(Kern kern, const DecArgs& args, size_t smem, cudaStream_t stream, int tile_points = kDecTile,
                      bool cooperative = false) {
    const int64_t num_tiles = (args.n + tile_points - 1) / tile_points;
    int64_t grid = persistent_grid(reinterpret_cast<const void*>(kern), kDecThreads, smem, kMaxCtasPerSm);
    if (grid < 0) return (int)grid;
    if (grid > kMaxDecGrid) grid = kMaxDecGrid;
    if (grid > num_tiles) grid = num_tiles;
    if (args.batch > 0 && grid > kBatchGridX) grid = kBatchGridX;    // R restarts share the machine (workspace bound)
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid, args.batch > 0 ? (unsigned)args.batch : 1u);
    cfg.blockDim = dim3(kDecThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (cooperative) {              // grid-wide barrier inside: every CTA must be resident (grid <= SMs x occupancy)
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
    } else {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // PDL, see scc_common.cuh
        attr[0].val.programmaticStreamSerializationAllowed = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SCC_CUDA(cudaLaunchKernelEx(&cfg, kern, args));
    return SCC_OK;
}

template <int D, int KP>
struct DecOps {
    static int assign(const DecArgs& a, cudaStream_t st) {
        const bool exact = a.K == KP, a1 = a.alpha == 1.0f;
        constexpr size_t smem = assign_smem<D, KP>();
        constexpr int tp = kDecTile * assign_ppt<D, KP>();
        if (exact && a1) return launch_dec(dec_assign_kernel<D, KP, true, true>, a, smem, st, tp);
        if (exact) return launch_dec(dec_assign_kernel<D, KP, true, false>, a, smem, st, tp);
        if (a1) return launch_dec(dec_assign_kernel<D, KP, false, true>, a, smem, st, tp);
        return launch_dec(dec_assign_kernel<D, KP, false, false>, a, smem, st, tp);
    }
    template <int MODE, bool EXACT, bool A1>
    static int grad_inst(const DecArgs& a, cudaStream_t st) {
        constexpr bool kTiled = (KP * D > 160);
        if constexpr (kTiled) {
            if constexpr (MODE == MODE_STEP)
                return SCC_ERR_UNSUPPORTED;          // the one-kernel step exists for the register-blocked shapes only
            else if constexpr (D % 4 == 0 && KP % 4 == 0)
                return launch_dec(dec_grad_tiled_kernel<D, KP, EXACT, A1, MODE>, a, grad_tiled_smem<D, KP>(), st);
            else
                return SCC_ERR_UNSUPPORTED;
        } else {
            return launch_dec(dec_grad_reg_kernel<D, KP, EXACT, A1, MODE>, a, grad_reg_smem<D, KP>(), st, kDecTile,
                              MODE == MODE_STEP);
        }
    }
    template <int MODE>
    static int grad(const DecArgs& a, cudaStream_t st) {
        const bool exact = a.K == KP, a1 = a.alpha == 1.0f;
        if (exact && a1) return grad_inst<MODE, true, true>(a, st);
        if (exact) return grad_inst<MODE, true, false>(a, st);
        if (a1) return grad_inst<MODE, false, true>(a, st);
        return grad_inst<MODE, false, false>(a, st);
    }
};

}  // namespace scc
