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

#include <cstdlib>
#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

constexpr int kDecThreads = 256;
constexpr int kDecTile = 256;
constexpr int kBatchGridX = 296;       // grid.x bound of a batched (grid.y = restarts) Lloyd launch
constexpr int kFixOffset = 8;          // workspace header: u64[8 .. 8+K] = fixed-point f accumulators of the one-kernel step

// doubles of reduction scratch for an NV-long statistics vector: cta_reduce needs
// [num_warps][round_up(NV, 32)], grid_publish needs 2 * kDecThreads
__host__ __device__ constexpr int reduce_scratch(int nv, int nt = kDecThreads) {
    return (nt / 32) * ((nv + 31) / 32 * 32) > 2 * nt ? (nt / 32) * ((nv + 31) / 32 * 32) : 2 * nt;
}

template <int D>
__host__ __device__ constexpr int dec_stages() { return RowLayout<D>::kDense ? 4 : (RowLayout<D>::kVec4 ? 3 : 2); }

// ---------------------------------------------------------------------------
// Packed FP32 (sm_100 FFMA2 / FADD2 / FMUL2): two lanes of a float2 per instruction.  On B200 a
// stream of 3-register scalar FFMAs issues at ~56 % of the FP32 peak (register-read bandwidth)
// while FFMA2 reaches ~88 % (tools/ubench_fp32.cu), and the issue-slot count halves.
//
// The two lanes of a pair are two CLUSTERS (2jp, 2jp+1), not two dimensions: every per-cluster
// quantity of a point (squared distance, w, u, t, q, p, gradient coefficient) then lives in
// KP/2 float2 registers and the whole per-cluster chain — not only the distance loop — runs on
// packed instructions, with the point's coordinate / the row-wide scalars riding in the .F32
// broadcast operand of FFMA2/FADD2/FMUL2.  Per lane the operations and their order are exactly
// those of a scalar evaluation: distance accumulated over c = 0..D-1 in ONE fma chain, so every
// kernel that evaluates q this way produces bit-identical values.
// ---------------------------------------------------------------------------
template <int D>
struct Pairs { static constexpr int N = (D + 1) / 2; };

__device__ __forceinline__ float2 pair_sel(bool cx, bool cy, float2 v, float other) {
    return make_float2(cx ? v.x : other, cy ? v.y : other);
}

__device__ __forceinline__ float2 round_dec5_2(float2 x) {
    // np.round(x, 5) on both lanes, see round_dec5()
    const float2 y = __ffma2_rn(x, make_float2(100000.0f, 100000.0f), make_float2(12582912.0f, 12582912.0f));
    return __fmul2_rn(__fadd2_rn(y, make_float2(-12582912.0f, -12582912.0f)), make_float2(1.0e-5f, 1.0e-5f));
}

// negated centroids, transposed and paired over clusters: nmuT2[c][jp] = -(mu[2jp][c], mu[2jp+1][c]);
// clusters j >= K are 0.  [D][KP/2] float2, 16-byte aligned.
template <int D, int KP>
__device__ __forceinline__ void load_neg_centroid_pairs(const float* __restrict__ mu, int K, float2* nmuT2) {
    float* flat = reinterpret_cast<float*>(nmuT2);
    for (int i = threadIdx.x; i < D * KP; i += blockDim.x) {
        const int c = i / KP, j = i - c * KP;
        flat[i] = (j < K) ? -mu[j * D + c] : 0.f;
    }
}

// ---------------------------------------------------------------------------
// Squared distances of P points (one thread) to every centroid, P points sharing each centroid load:
// acc2[r][jp] = (||z_r - mu_2jp||^2, ||z_r - mu_2jp+1||^2).  networks.py:280-282.
// ---------------------------------------------------------------------------
template <int D, int KP, int P>
__device__ __forceinline__ void sq_distances(const float (&z)[P][D], const float2* __restrict__ nmuT2,
                                             float2 (&acc2)[P][KP / 2]) {
    constexpr int JP = KP / 2;
#pragma unroll
    for (int r = 0; r < P; ++r)
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) acc2[r][jp] = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < D; ++c) {
        const float4* row = reinterpret_cast<const float4*>(nmuT2 + c * JP);
#pragma unroll
        for (int h = 0; h < JP / 2; ++h) {
            const float4 m = row[h];
#pragma unroll
            for (int r = 0; r < P; ++r) {
                const float2 zc = make_float2(z[r][c], z[r][c]);
                const float2 d0 = __fadd2_rn(zc, make_float2(m.x, m.y));
                const float2 d1 = __fadd2_rn(zc, make_float2(m.z, m.w));
                acc2[r][2 * h] = __ffma2_rn(d0, d0, acc2[r][2 * h]);
                acc2[r][2 * h + 1] = __ffma2_rn(d1, d1, acc2[r][2 * h + 1]);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Student's-t kernel of one point from its squared distances.  networks.py:283-287, models.py:92.
//   w_j = 1 + d_j / alpha,  u_j = 1 / w_j,  t_j = u_j^((alpha+1)/2),  tsum = sum_j t_j,  q_j = t_j / tsum.
// LABEL: also the hard label = argmin distance (== argmax q, first index wins) and that distance.
// Reciprocals are single MUFU.RCP instructions (w >= 1, so no range fix-up is needed).
// ---------------------------------------------------------------------------
template <int KP, bool EXACT, bool ALPHA1, bool LABEL>
__device__ __forceinline__ void student_t_pairs(const float2 (&acc2)[KP / 2], int K, float inv_alpha, float expo,
                                                float2 (&w2)[KP / 2], float2 (&u2)[KP / 2], float2 (&t2)[KP / 2],
                                                float& tsum, int& label, float& best) {
    constexpr int JP = KP / 2;
    float2 ts2 = make_float2(0.f, 0.f);
    best = 3.4e38f;
    label = 0;
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) {
        const bool vx = EXACT || 2 * jp < K, vy = EXACT || 2 * jp + 1 < K;
        if (LABEL) {
            if (vx && acc2[jp].x < best) { best = acc2[jp].x; label = 2 * jp; }
            if (vy && acc2[jp].y < best) { best = acc2[jp].y; label = 2 * jp + 1; }
        }
        float2 ww = ALPHA1 ? __fadd2_rn(acc2[jp], make_float2(1.f, 1.f))
                           : __ffma2_rn(acc2[jp], make_float2(inv_alpha, inv_alpha), make_float2(1.f, 1.f));
        float2 uu = make_float2(rcp_approx(ww.x), rcp_approx(ww.y));
        float2 tt = ALPHA1 ? uu : make_float2(ex2_approx(-expo * lg2_approx(ww.x)), ex2_approx(-expo * lg2_approx(ww.y)));
        if (!EXACT) {
            ww = pair_sel(vx, vy, ww, 1.f); uu = pair_sel(vx, vy, uu, 0.f); tt = pair_sel(vx, vy, tt, 0.f);
        }
        w2[jp] = ww; u2[jp] = uu; t2[jp] = tt;
        ts2 = __fadd2_rn(ts2, tt);
    }
    tsum = ts2.x + ts2.y;
}

template <int KP, bool EXACT>
__device__ __forceinline__ void store_krow2(float* __restrict__ dst, int K, const float2 (&v)[KP / 2]) {
    if (EXACT || (K & 3) == 0) {
#pragma unroll
        for (int j = 0; j < KP; j += 4)
            if (EXACT || j < K)
                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j / 2].x, v[j / 2].y, v[j / 2 + 1].x, v[j / 2 + 1].y);
    } else {
#pragma unroll
        for (int j = 0; j < KP; ++j)
            if (j < K) dst[j] = (j & 1) ? v[j / 2].y : v[j / 2].x;
    }
}
template <int KP, bool EXACT>
__device__ __forceinline__ void load_krow2(const float* __restrict__ src, int K, float2 (&v)[KP / 2]) {
    if (EXACT || (K & 3) == 0) {
#pragma unroll
        for (int j = 0; j < KP; j += 4) {
            if (EXACT || j < K) {
                const float4 x = ldg_stream4(reinterpret_cast<const float4*>(src + j));
                v[j / 2] = make_float2(x.x, x.y); v[j / 2 + 1] = make_float2(x.z, x.w);
            } else { v[j / 2] = make_float2(0.f, 0.f); v[j / 2 + 1] = make_float2(0.f, 0.f); }
        }
    } else {
#pragma unroll
        for (int jp = 0; jp < KP / 2; ++jp)
            v[jp] = make_float2((2 * jp < K) ? ldg_stream(src + 2 * jp) : 0.f, (2 * jp + 1 < K) ? ldg_stream(src + 2 * jp + 1) : 0.f);
    }
}
// scalar row load kept for colsum_kernel (dec_api.cu)
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
// Soft assignment of P rows per thread (assign pass): q2[r][jp], label[r].  The centroid loads (one
// LDS.128 per two cluster pairs and dimension) are shared by the P points: they are what saturates
// first at large K*d (LSU pipe: 1 wavefront/clk/SM against 4 FP32 warp-instr/clk).
// ---------------------------------------------------------------------------
template <int D, int KP, bool EXACT, bool ALPHA1, int P>
__device__ __forceinline__ void soft_assign_rows(const float (&z)[P][D], const float2* __restrict__ nmuT2,
                                                 int K, float inv_alpha, float expo, bool round5,
                                                 float2 (&q2)[P][KP / 2], int (&label)[P],
                                                 float* const (&u_out)[P]) {
    constexpr int JP = KP / 2;
    float2 acc2[P][JP];
    sq_distances<D, KP, P>(z, nmuT2, acc2);
#pragma unroll
    for (int r = 0; r < P; ++r) {
        float2 w2[JP], u2[JP];
        float tsum, best;
        student_t_pairs<KP, EXACT, ALPHA1, true>(acc2[r], K, inv_alpha, expo, w2, u2, q2[r], tsum, label[r], best);
        if (u_out[r]) store_krow2<KP, EXACT>(u_out[r], K, u2);      // hand-off to the gradient pass (MODE_KLU)
        const float inv = rcp_approx(tsum);
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) {
            q2[r][jp] = __fmul2_rn(q2[r][jp], make_float2(inv, inv));
            if (round5) q2[r][jp] = round_dec5_2(q2[r][jp]);
        }
    }
}

// points per thread of the assign pass (tile = 256 * P points)
template <int D, int KP>
__host__ __device__ constexpr int assign_ppt() {
#ifdef SCC_ASSIGN_PPT2
    return 2;                           // A/B builds
#else
    return (KP * D <= 96) ? 2 : 1;
#endif
}
template <int D, int KP>
__host__ __device__ constexpr int assign_stages() {
    // stages x 256 * P rows must leave room for two CTAs per SM where the registers allow them
    return RowLayout<D>::kDense ? (assign_ppt<D, KP>() == 2 ? 3 : 4)
                                : (RowLayout<D>::kVec4 ? ((assign_ppt<D, KP>() == 2 && RowLayout<D>::LD >= 28) ? 2 : 3) : 2);
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
    constexpr int JP = KP / 2;
    // CTA-level ring: a per-warp ring was measured for this kernel too — its 6x more, 6x smaller
    // TMA copies lengthen the prologue by ~1.3 us and the short main loop gains nothing
    using Ring = ZRing<D, TILE, S, kDecThreads>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float2* nmuT2 = reinterpret_cast<float2*>(ring_buf + ((S * Ring::kTileFloats + 3) & ~3));     // [D][JP] (-mu pairs)
    double* cta_stats = reinterpret_cast<double*>(nmuT2 + ((D * JP + 1) & ~1));        // [KP+1]
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
    load_neg_centroid_pairs<D, KP>(a.mu, K, nmuT2);
    __syncthreads();
    SCC_TL(a.timeline, 1);

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    const bool round5 = a.round5 != 0;
    float2 facc2[JP];
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) facc2[jp] = make_float2(0.f, 0.f);
    float changed = 0.f;

    int stage = 0;
    uint32_t use = 0;
    for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G) {
        ring.wait(stage, tile, use);
        if (tile == (int)blockIdx.x) SCC_TL(a.timeline, 2);
        const int np = ring.points(tile);
        float zr[P][D];
        bool active[P];
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const int t = threadIdx.x + r * kDecThreads;
            active[r] = t < np;
#pragma unroll
            for (int c = 0; c < D; ++c) zr[r][c] = 0.f;
            if (active[r]) load_row<D>(ring.stage_ptr(stage), t, zr[r]);
        }
        __syncthreads();                 // every row of the stage is in registers: refill it
        ring.issue(stage, tile + S * G);
        if (active[0]) {
            float2 q2[P][JP];
            int label[P];
            float* up[P];
#pragma unroll
            for (int r = 0; r < P; ++r)
                up[r] = (a.u_out && active[r]) ? a.u_out + ((size_t)tile * TILE + threadIdx.x + r * kDecThreads) * K : nullptr;
            soft_assign_rows<D, KP, EXACT, ALPHA1, P>(zr, nmuT2, K, inv_alpha, expo, round5, q2, label, up);
#pragma unroll
            for (int r = 0; r < P; ++r) {
                if (active[r]) {
                    const size_t i = (size_t)tile * TILE + threadIdx.x + r * kDecThreads;
#pragma unroll
                    for (int jp = 0; jp < JP; ++jp) facc2[jp] = __fadd2_rn(facc2[jp], q2[r][jp]);
                    if (a.q) store_krow2<KP, EXACT>(a.q + i * K, K, q2[r]);
                    if (a.labels) a.labels[i] = label[r];
                    if (a.labels_prev) changed += (a.labels_prev[i] != label[r]) ? 1.f : 0.f;
                }
            }
        }
        if (++stage == S) { stage = 0; ++use; }
    }
    pdl_trigger();                      // successor may start its prologue under our reduction tail
    SCC_TL(a.timeline, 3);
    __syncthreads();                    // every warp is done with the ring: it becomes the reduction scratch
    float facc[KP + 1];
#pragma unroll
    for (int j = 0; j < KP; ++j) facc[j] = (j & 1) ? facc2[j / 2].y : facc2[j / 2].x;
    facc[KP] = changed;
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
// Inputs are the Student's-t quantities of student_t_pairs(): q_j = t_j / tsum, 1/q_j = tsum w_j^expo.
// Everything per cluster is a float2 over the cluster pair (2jp, 2jp+1).
// ---------------------------------------------------------------------------
// The [n, K] operand a gradient kernel streams besides z: the target p (MODE_KL, API mode) or the
// upstream gradient dL/dq (MODE_GENERIC).
template <int MODE>
__device__ __forceinline__ const float* krow_operand(const DecArgs& a) {
    return MODE == MODE_KL ? a.p : (MODE == MODE_GENERIC ? a.grad_q : nullptr);
}

template <int MODE>
__host__ __device__ constexpr bool mode_is_kl() { return MODE == MODE_KL || MODE == MODE_KLF || MODE == MODE_STEP || MODE == MODE_KLU; }

template <int MODE>
__host__ __device__ __forceinline__ float grad_fold_scale(float scale, float alpha) {
    return MODE == MODE_KMEANS ? 1.f : (mode_is_kl<MODE>() ? scale : 1.f) * (alpha + 1.f) / alpha;
}

template <int KP, bool EXACT, bool ALPHA1, int MODE>
__device__ __forceinline__ void grad_coefficients(const DecArgs& a, size_t i, int K, const float2* __restrict__ inv_f2,
                                                  const float2 (&w2)[KP / 2], const float2 (&u2)[KP / 2],
                                                  const float2 (&t2)[KP / 2], float tsum, float expo, int label,
                                                  float best, const float2 (&pre2)[KP / 2],
                                                  float2 (&coef2)[KP / 2], float& loss, float& ssum) {
    constexpr int JP = KP / 2;
    if constexpr (MODE == MODE_KMEANS) {
        // Lloyd statistics: one-hot coefficient on the nearest centre, "loss" = inertia
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) coef2[jp] = make_float2(2 * jp == label ? 1.f : 0.f, 2 * jp + 1 == label ? 1.f : 0.f);
        loss += best;
        if (a.labels) a.labels[i] = label;
        if (a.mindist) a.mindist[i] = best;
    } else if constexpr (mode_is_kl<MODE>()) {
        const float inv = rcp_approx(tsum);
        float2 p2[JP];
        if constexpr (MODE == MODE_KL) {           // target row streamed from memory
#pragma unroll
            for (int jp = 0; jp < JP; ++jp) p2[jp] = pre2[jp];
        } else {                                   // MODE_KLF / MODE_STEP / MODE_KLU: rebuild p from the column sums
            const float2 inv2 = make_float2(inv, inv);
            float2 qrow[JP];
#pragma unroll
            for (int jp = 0; jp < JP; ++jp) {
                float2 q = __fmul2_rn(t2[jp], inv2);
                if (a.round5) q = round_dec5_2(q);
                if (MODE == MODE_STEP) qrow[jp] = q;
                p2[jp] = __fmul2_rn(__fmul2_rn(q, q), inv_f2[jp]);         // inv_f is 0 for clusters >= K
            }
            // one-kernel step: q leaves from the FP32-bound second pass (the same value, bit for bit, as the assign
            // pass computes), which takes 4K bytes per point off the HBM-bound first pass (47.2 -> 45.8 us at 1M points)
            if (MODE == MODE_STEP && a.q) store_krow2<KP, EXACT>(a.q + i * K, K, qrow);
            // row sum in the order dec_target_kernel uses (groups of 4, then a pairwise tree over the
            // groups; sequential when K is not 4, 8 or 16), so the rebuilt p is bit-identical to its output
            float wsum;
            if ((K & 3) == 0 && K != 12) {
                float g4[KP / 4];
#pragma unroll
                for (int b = 0; b < KP / 4; ++b) g4[b] = (p2[2 * b].x + p2[2 * b].y) + (p2[2 * b + 1].x + p2[2 * b + 1].y);
                if constexpr (KP == 4) wsum = g4[0];
                else if constexpr (KP == 8) wsum = (K == 4) ? g4[0] : g4[0] + g4[1];
                else wsum = (K == 4) ? g4[0] : ((K == 8) ? g4[0] + g4[1] : (g4[0] + g4[1]) + (g4[2] + g4[3]));
            } else {
                wsum = 0.f;
#pragma unroll
                for (int j = 0; j < KP; ++j) wsum += (EXACT || j < K) ? ((j & 1) ? p2[j / 2].y : p2[j / 2].x) : 0.f;
            }
            const float winv = 1.f / wsum;
            const float2 winv2 = make_float2(winv, winv);
#pragma unroll
            for (int jp = 0; jp < JP; ++jp) {
                p2[jp] = __fmul2_rn(p2[jp], winv2);
                if (a.round5) p2[jp] = round_dec5_2(p2[jp]);
            }
            if (a.p_out) store_krow2<KP, EXACT>(a.p_out + i * K, K, p2);      // materialise target_distribution(q)
        }
        // p_j / q_j = p_j w_j tsum for alpha == 1 (w_j = 1 + d_j is already in registers: no division); the log is
        // taken of the RATIO (near 1), not of its large factors separately — lg2.approx has a relative error, so
        // log2(p w) + log2(tsum) would lose the digits that cancel.  The 1e-37 keeps a zero target at
        // 0 * finite = 0 (torch KLDivLoss: xlogy); negative / NaN targets still give NaN.
        float2 s2 = make_float2(0.f, 0.f), l2 = make_float2(0.f, 0.f);
        const float2 tiny2 = make_float2(1e-37f, 1e-37f);
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) {
            s2 = __fadd2_rn(s2, p2[jp]);
            float2 ratio;
            if (ALPHA1) {
                ratio = __ffma2_rn(__fmul2_rn(p2[jp], w2[jp]), make_float2(tsum, tsum), tiny2);
            } else {
                ratio = make_float2(fmaf(p2[jp].x, rcp_approx(fmaxf(t2[jp].x * inv, 1e-37f)), 1e-37f),
                                    fmaf(p2[jp].y, rcp_approx(fmaxf(t2[jp].y * inv, 1e-37f)), 1e-37f));
            }
            l2 = __ffma2_rn(p2[jp], make_float2(lg2_approx(ratio.x), lg2_approx(ratio.y)), l2);
        }
        const float s = s2.x + s2.y;
        const float nis = -(inv * s);
        const float2 nis2 = make_float2(nis, nis);
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) coef2[jp] = __fmul2_rn(__ffma2_rn(t2[jp], nis2, p2[jp]), u2[jp]);
        loss += l2.x + l2.y;
        ssum += s;
    } else {
        const float inv = rcp_approx(tsum);
        float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) d2 = __ffma2_rn(pre2[jp], t2[jp], d2);
        const float dot = (d2.x + d2.y) * inv;
        const float2 inv2 = make_float2(inv, inv), dot2 = make_float2(dot, dot);
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) {
            const float2 ng = make_float2(-pre2[jp].x, -pre2[jp].y);
            coef2[jp] = __fmul2_rn(__fmul2_rn(__fmul2_rn(t2[jp], inv2), __fadd2_rn(dot2, ng)), u2[jp]);
        }
    }
}

// dz_c = cs ((sum_j c_j) zc_c - sum_j c_j mc_jc)  with zc = z - c0, mc = mu - c0 (algebraic form of
// sum_j c_j (z_c - mu_jc): K*D/2 FFMA2 instead of K*D (FADD + FMA)).  nmc2_s = -cs (mu - c0) as pairs over
// the DIMENSION [KP][ceil(D/2)], csum = cs sum_j c_j.  The pad lane of an odd D holds garbage and is never stored.
template <int D, int KP, bool EXACT>
__device__ __forceinline__ void dz_from_coefficients(const float2 (&zc2)[Pairs<D>::N], const float2 (&coef2)[KP / 2],
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
            const float2 cj = splat2((j & 1) ? coef2[j / 2].y : coef2[j / 2].x);
#pragma unroll
            for (int c = 0; c < DP2; ++c) dz2[c] = __ffma2_rn(cj, nmc2_s[j * DP2 + c], dz2[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < D; ++c) dzr[c] = (c & 1) ? dz2[c >> 1].y : dz2[c >> 1].x;
}

// The same for RB rows of one thread sharing every load of the -(mu - c0) table.
template <int D, int KP, bool EXACT, int RB>
__device__ __forceinline__ void dz_from_coefficients_rows(const float2 (&zc2)[RB][Pairs<D>::N],
                                                          const float2 (&coef2)[RB][KP / 2], const float (&csum)[RB],
                                                          const float2* __restrict__ nmc2_s, int K,
                                                          float (&dzr)[RB][D]) {
    constexpr int DP2 = Pairs<D>::N;
    float2 dz2[RB][DP2];
#pragma unroll
    for (int k = 0; k < RB; ++k)
#pragma unroll
        for (int c = 0; c < DP2; ++c) dz2[k][c] = __fmul2_rn(splat2(csum[k]), zc2[k][c]);
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        if (EXACT || j < K) {
#pragma unroll
            for (int c = 0; c < DP2; ++c) {
                const float2 m = nmc2_s[j * DP2 + c];
#pragma unroll
                for (int k = 0; k < RB; ++k)
                    dz2[k][c] = __ffma2_rn(splat2((j & 1) ? coef2[k][j / 2].y : coef2[k][j / 2].x), m, dz2[k][c]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < RB; ++k)
#pragma unroll
        for (int c = 0; c < D; ++c) dzr[k][c] = (c & 1) ? dz2[k][c >> 1].y : dz2[k][c >> 1].x;
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

// Shared prologue of the gradient kernels: -mu cluster pairs (transposed), -cs (mu - c0) dimension pairs,
// c0, (mu - c0), 1/f.
template <int D, int KP>
__device__ __forceinline__ void load_grad_constants(const DecArgs& a, int K, float cs, float2* nmuT2,
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
    load_neg_centroid_pairs<D, KP>(a.mu, K, nmuT2);
    float* nmc = reinterpret_cast<float*>(nmc2_s);
    for (int i = threadIdx.x; i < KP * DP2 * 2; i += blockDim.x) {
        const int j = i / (2 * DP2), c = i - j * (2 * DP2);
        const bool ok = (j < K && c < D);
        nmc[i] = ok ? -cs * (a.mu[j * D + c] - c0_s[c]) : 0.f;
    }
    for (int i = threadIdx.x; i < KP * D; i += blockDim.x) mc_s[i] = (i < K * D) ? a.mu[i] - c0_s[i % D] : 0.f;
}

// ---------------------------------------------------------------------------
// Per-warp stream of latent rows (register-blocked gradient kernel / one-kernel step).
//   The rows are split into contiguous blocks, one per WARP of the grid (blocks differ by at most one
//   32-row slice: a static, deterministic, balanced partition whatever the stage size); a warp walks its
//   block in stages of up to 32*P rows — ONE TMA bulk copy onto the warp's own mbarrier (padded layouts:
//   the warp's lanes issue cp.async / plain loads) — and the only synchronisation in the main loop is
//   __syncwarp().  Lane l owns rows l, l+32, ... of a stage and processes them one after the other, so the
//   ring / barrier / copy-out bookkeeping is paid once per P rows of every thread.
//   dL/dz rows are written IN PLACE over the consumed z rows and leave with one bulk store per stage; the
//   stage is refilled after the first slice of the NEXT stage has been processed (by then the store has
//   long finished reading shared memory, and the refill still has P-1 slices of compute to land behind).
// ---------------------------------------------------------------------------
template <int D, int P, int STAGES>
struct WarpStream {
    using L = RowLayout<D>;
    static constexpr int kRows = 32 * P;
    static constexpr int kStageFloats = kRows * L::LD;
    static constexpr bool kCpAsync = L::kVec4 && !L::kDense;
    static_assert(STAGES >= 2, "stream needs two stages");

    float* buf;            // this warp's STAGES stages
    uint64_t* bar;         // this warp's STAGES mbarriers
    const float* zw;       // first row of this warp's block
    int64_t row_begin;     // its global row index
    int total, req;        // rows in the block / rows requested so far (a warp's block is far below 2^31 rows)
    int lane;
    uint32_t phase;        // bit s = parity the next wait on stage s expects (flips per completed TMA fill)

    __device__ __forceinline__ void init(float* warp_buf, uint64_t* warp_bars, const float* z_, int64_t n,
                                         int warp_global, int warps_total) {
        buf = warp_buf; bar = warp_bars; phase = 0u; lane = threadIdx.x & 31;
        const int64_t slices = (n + 31) >> 5;
        const int64_t base = slices / warps_total, rem = slices - base * warps_total;
        const int64_t s0 = warp_global * base + (warp_global < rem ? warp_global : rem);
        const int64_t cnt = base + (warp_global < rem ? 1 : 0);
        row_begin = 32 * s0 < n ? 32 * s0 : n;
        const int64_t row_end = 32 * (s0 + cnt) < n ? 32 * (s0 + cnt) : n;
        total = (int)(row_end - row_begin);
        req = 0;
        zw = z_ + (size_t)row_begin * D;
        if (L::kDense && lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    __device__ __forceinline__ void rewind() { req = 0; }
    __device__ __forceinline__ float* stage_ptr(int stage) const { return buf + stage * kStageFloats; }
    static __device__ __forceinline__ bool tma_ok(int rows) { return L::kDense && rows > 0 && ((rows * D) & 3) == 0; }
    // rows of the stage that starts at local row `from`
    __device__ __forceinline__ int stage_rows(int from) const {
        const int left = total - from;
        return left < kRows ? left : kRows;
    }

    // Called by all lanes of the warp (converged): request the next stage of this warp's block.
    __device__ __forceinline__ void issue(int stage) {
        const int rows = stage_rows(req);
        if (rows > 0) {
            float* dst = stage_ptr(stage);
            const float* src = zw + (size_t)req * D;
            if (tma_ok(rows)) {
                if (lane == 0) {
                    mbar_expect_tx(&bar[stage], (uint32_t)rows * D * sizeof(float));
                    bulk_g2s(dst, src, (uint32_t)rows * D * sizeof(float), &bar[stage]);
                }
            } else if constexpr (L::kVec4) {
                const int nvec = rows * (D / 4);
                const float4* src4 = reinterpret_cast<const float4*>(src);
                for (int v = lane; v < nvec; v += 32) {
                    const int row = v / (D / 4), c4 = v - row * (D / 4);
                    if constexpr (kCpAsync) cp_async16(dst + row * L::LD + 4 * c4, src4 + v);
                    else *reinterpret_cast<float4*>(dst + row * L::LD + 4 * c4) = ldg_stream4(src4 + v);
                }
            } else {
                const int nf = rows * D;
                for (int f = lane; f < nf; f += 32) {
                    const int row = f / D, c = f - row * D;
                    dst[row * L::LD + c] = ldg_stream(src + f);
                }
            }
            req += rows;
        }
        if constexpr (kCpAsync) cp_async_commit();        // empty groups keep the per-thread count aligned
    }
    // After wait() returns every lane may read any row of the stage.  The awaited fill is the youngest
    // group but STAGES-2 (a stage is refilled one stage late, see above).  Stages filled with plain stores
    // were written at least one __syncwarp() ago.
    __device__ __forceinline__ void wait(int stage, int rows) {
        if constexpr (kCpAsync) {
            cp_async_wait<STAGES - 2>();
            __syncwarp();
        } else if constexpr (L::kDense) {
            if (tma_ok(rows)) {
                mbar_wait(&bar[stage], (phase >> stage) & 1u);
                phase ^= 1u << stage;
            }
        }
    }
};

// rows per thread and stage of the register-blocked kernels: 2 stages x 8 warps x 32P rows x LD floats <= ~80 KB
template <int D>
__host__ __device__ constexpr int reg_ppt() {
#ifdef SCC_REG_PPT
    return SCC_REG_PPT;                 // A/B builds (make variant VFLAGS=-DSCC_REG_PPT=2)
#else
    return 40 / RowLayout<D>::LD > 4 ? 4 : (40 / RowLayout<D>::LD < 1 ? 1 : 40 / RowLayout<D>::LD);
#endif
}
#ifdef SCC_REG_STAGES
constexpr int kRegStages = SCC_REG_STAGES;     // A/B builds
#else
constexpr int kRegStages = 2;
#endif
// rows a thread of the register-blocked kernels holds at once: the broadcast operand loads (centroid table of
// the distance loop, -(mu - c0) table of dz) are shared by the RB rows.  A warp-wide LDS.128 occupies the
// shared-memory return path for 4 cycles whatever the broadcast, so with one row at a time those loads alone
// are ~150 LSU cycles per 32-row slice and pass (d = 9, K = 8).
// Measured (profiles/r02_variants.txt, 16M points, d = 9, K = 8): two rows at once pay where a second [n, K]
// operand is streamed from HBM (dec_kl_grad(p): 439 -> 355 us, the rows' p loads overlap) and cost where the row's
// own chain is the limit (one-kernel step 560 -> 598 us: the operand loads were not the bound, the extra live
// registers spill) — so the block is 2 rows for MODE_KL / MODE_GENERIC and 1 row otherwise.
template <int MODE>
__host__ __device__ constexpr int reg_row_block() {
#ifdef SCC_REG_RB
    return SCC_REG_RB;                     // A/B builds
#else
    return (MODE == MODE_KL || MODE == MODE_GENERIC) ? 2 : 1;
#endif
}

// Copy `rows` staged rows (row stride LD) of this warp to rows * D contiguous floats at dst, all lanes.
template <int D>
__device__ __forceinline__ void warp_copy_rows_out(const float* __restrict__ src, float* __restrict__ dst, int rows) {
    using L = RowLayout<D>;
    const int lane = threadIdx.x & 31;
    if constexpr (L::kVec4) {
        const int nvec = rows * (D / 4);
        for (int v = lane; v < nvec; v += 32) {
            const int row = v / (D / 4), c4 = v - row * (D / 4);
            reinterpret_cast<float4*>(dst)[v] = *reinterpret_cast<const float4*>(src + row * L::LD + 4 * c4);
        }
    } else {
        const int nf = rows * D;
        for (int f = lane; f < nf; f += 32) {
            const int row = f / D, c = f - row * D;
            dst[f] = src[row * L::LD + c];
        }
    }
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
// dec_grad, REG variant.  Per-thread accumulators (float2 over cluster pairs):
//   B_jc = sum_i c_ij (z_ic - c0_c),  W_j = sum_i c_ij;   dmu_jc = -cs (B_jc - W_j (mu_jc - c0_c)).
// stats out: [loss, sum_i s_i, dmu[K*D]]
// MODE_STEP: the same persistent warps first run the assign pass over their rows, meet at a grid-wide barrier
// that all-reduces f (over the CTAs and, multi-GPU, over NVLink), then run the target + gradient pass.
// ---------------------------------------------------------------------------
template <int D, int KP>
__host__ __device__ constexpr int grad_reg_accumulators() { return 2 + KP * (D + 1); }
// Threads per CTA of the register-blocked gradient / step kernel: ONE CTA per SM.
//   * K*(d+1) + 2 <= 100 accumulators (d = 9, K = 8: 82): 384 threads = 12 warps at 168 registers — the accumulators
//     plus the per-cluster chain fit without spilling (at the 128-register cap of 16 warps the kernel carried ~40
//     local-memory accesses per point: one-kernel step 57.3 vs 51.3 us at 1M points).  One CTA of 12 warps instead
//     of two of 6 halves the per-CTA slots of the grid reduction and shares one set of constant tables:
//     step 44.7 -> 43.5 us at 1M, 562 -> 537 us at 16M; dec_target_kl_grad 369 -> 340 us at 16M (r02_variants.txt).
//   * more accumulators (d = 9, K = 16: 162): 192 threads at up to 255 registers.
template <int D, int KP>
__host__ __device__ constexpr int reg_threads() {
#ifdef SCC_REG_THREADS
    return SCC_REG_THREADS;                    // A/B builds (make variant VFLAGS="-DSCC_REG_THREADS=192 -DSCC_REG_CTAS=2")
#else
    return grad_reg_accumulators<D, KP>() <= 100 ? 384 : 192;
#endif
}
template <int D, int KP>
__host__ __device__ constexpr int grad_reg_ctas_per_sm() {
#ifdef SCC_REG_CTAS
    return SCC_REG_CTAS;
#else
    return 1;
#endif
}
template <int D, int KP>
__host__ __device__ constexpr size_t grad_reg_smem() {
    constexpr int S = kRegStages, P = reg_ppt<D>(), NT = reg_threads<D, KP>(), NW = NT / 32;
    constexpr int NV = 2 + KP + KP * D;
    constexpr int SCR = reduce_scratch(NV, NT);
    return sizeof(float) * (((NW * S * 32 * P * RowLayout<D>::LD + 3) & ~3) + 2 * ((D * (KP / 2) + 1) & ~1) +
                            2 * ((KP * Pairs<D>::N + 1) & ~1) +
                            ((KP * D + 3) & ~3) + ((D + 3) & ~3) + ((KP + 3) & ~3) + 2 * ((Pairs<D>::N + 1) & ~1)) +
           sizeof(double) * (SCR + NV) + sizeof(uint64_t) * S * NW;
}

template <int D, int KP, bool EXACT, bool ALPHA1, int MODE>
__global__ void __launch_bounds__((reg_threads<D, KP>()), (grad_reg_ctas_per_sm<D, KP>()))
dec_grad_reg_kernel(const DecArgs a_in) {
    constexpr int S = kRegStages;
    constexpr int P = reg_ppt<D>();
    constexpr int RB = (P % reg_row_block<MODE>() == 0 && grad_reg_accumulators<D, KP>() <= 100)
                           ? reg_row_block<MODE>() : 1;                     // rows in registers at once
    constexpr int DP2 = Pairs<D>::N;
    constexpr int JP = KP / 2;
    constexpr int DW = D + 1;                                                // accumulator row: B_c (c < D), W
    constexpr int NT = reg_threads<D, KP>();
    constexpr int NW = NT / 32;
    using Stream = WarpStream<D, P, S>;
    using L = RowLayout<D>;
    constexpr int NV = 2 + KP + KP * D;                                      // [loss, sum s, W[KP], B[KP*D]]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);                    // [NW][S][32P*LD]
    float2* nmuT2 = reinterpret_cast<float2*>(ring_buf + ((NW * S * Stream::kStageFloats + 3) & ~3));   // [D][JP]
    float2* nmc2_s = nmuT2 + ((D * JP + 1) & ~1);                            // [KP][DP2]
    float* mc_s = reinterpret_cast<float*>(nmc2_s + ((KP * DP2 + 1) & ~1));  // [KP*D]
    float* c0_s = mc_s + ((KP * D + 3) & ~3);                                // [D]
    float* inv_f = c0_s + ((D + 3) & ~3);                                    // [KP] (read as cluster pairs)
    float2* nc0_s = reinterpret_cast<float2*>(inv_f + ((KP + 3) & ~3));      // [DP2] -c0 as dimension pairs
    double* scratch = reinterpret_cast<double*>(nc0_s + ((DP2 + 1) & ~1));   // [reduce_scratch(NV, NT)]
    double* cta_stats = scratch + reduce_scratch(NV, NT);                        // [NV]  (>= K*D + 2 + K)
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + NV);            // [NW][S]
    constexpr bool kBulkOut = L::kDense;

    const int K = EXACT ? KP : a_in.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Stream st;
    st.init(ring_buf + warp * S * Stream::kStageFloats, bars + warp * S, a_in.z, a_in.n,
            blockIdx.x * NW + warp, gridDim.x * NW);
    pdl_wait();                         // no global access before this point (see scc_common.cuh)
    DecArgs a_view = a_in;
    if (!batch_view<MODE, D>(a_view, K)) return;
    const DecArgs& a = (MODE == MODE_KMEANS) ? a_view : a_in;
    const float cs = grad_fold_scale<MODE>(a.scale, a.alpha);
    SCC_TL(a.timeline, 0);
#pragma unroll
    for (int s = 0; s < S; ++s) st.issue(s);
    __shared__ unsigned int s_seq1;             // sequence number of the f exchange (multi-GPU step): read here, made
    if (MODE == MODE_STEP && threadIdx.x == 0)  // visible by the barrier below, before any CTA can have advanced it
        s_seq1 = a.ex_windows ? ld_relaxed_gpu_u32(&reinterpret_cast<PeerHeader*>(a.ex_windows[a.ex_rank])->seq) + 1u : 0u;
    load_grad_constants<D, KP>(a, K, cs, nmuT2, nmc2_s, mc_s, c0_s, inv_f);
    if (threadIdx.x < DP2)
        nc0_s[threadIdx.x] = make_float2(-c0_s[2 * threadIdx.x], (2 * threadIdx.x + 1 < D) ? -c0_s[2 * threadIdx.x + 1] : 0.f);
    __syncthreads();
    SCC_TL(a.timeline, 1);

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    const bool want_dz = a.dz != nullptr;
    const float2* inv_f2 = reinterpret_cast<const float2*>(inv_f);

    if constexpr (MODE == MODE_STEP) {
        // ---- pass 1 of the one-kernel DEC step: the assign pass (same arithmetic as dec_assign_kernel) over
        // this warp's rows, then a grid-wide barrier + all-reduce of f, all CTAs co-resident (cooperative launch)
        float2 facc2[JP];
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) facc2[jp] = make_float2(0.f, 0.f);
        float changed = 0.f;
#ifdef SCC_STEP_RB1
        constexpr int RB1 = (P % SCC_STEP_RB1 == 0) ? SCC_STEP_RB1 : 1;      // A/B builds
#else
        constexpr int RB1 = 1;                  // rows of the assign pass in registers at once
#endif
        int cons = 0, stage = 0, pending = -1;
        while (cons < st.total) {
            const int rows = st.stage_rows(cons);
            st.wait(stage, rows);
            const float* sp = st.stage_ptr(stage);
            const int nsl = (rows + 31) >> 5;
            for (int r = 0; r < nsl; r += RB1) {
                if (32 * r + lane < rows) {             // row k of the block: slice r + k (rows ascend: k = 0 is valid)
                    float zr[RB1][D];
                    bool valid[RB1];
#pragma unroll
                    for (int k = 0; k < RB1; ++k) {
                        valid[k] = 32 * (r + k) + lane < rows;
#pragma unroll
                        for (int c = 0; c < D; ++c) zr[k][c] = 0.f;
                        if (valid[k]) load_row<D>(sp, 32 * (r + k) + lane, zr[k]);
                    }
                    float2 q2[RB1][JP];
                    int label[RB1];
                    float* no_u[RB1];
#pragma unroll
                    for (int k = 0; k < RB1; ++k) no_u[k] = nullptr;
                    soft_assign_rows<D, KP, EXACT, ALPHA1, RB1>(zr, nmuT2, K, inv_alpha, expo, a.round5 != 0, q2, label, no_u);
#pragma unroll
                    for (int k = 0; k < RB1; ++k) {
                        if (valid[k]) {
                            const size_t i = (size_t)st.row_begin + (cons + 32 * (r + k) + lane);
#pragma unroll
                            for (int jp = 0; jp < JP; ++jp) facc2[jp] = __fadd2_rn(facc2[jp], q2[k][jp]);
                            // (q itself is written by the second pass, see grad_coefficients())
                            if (a.labels) a.labels[i] = label[k];
                            if (a.labels_prev) changed += (a.labels_prev[i] != label[k]) ? 1.f : 0.f;
                        }
                    }
                }
                if (r == 0 && pending >= 0) {           // refill the stage consumed before this one
                    __syncwarp();
                    st.issue(pending);
                    pending = -1;
                }
            }
            __syncwarp();                               // every lane has its rows of this stage in registers
            cons += rows;
            pending = stage;
            if (++stage == S) stage = 0;
        }
        SCC_TL(a.timeline, 6);                                     // pass 1 main loop done
        // pass 2's first stages are requested before the grid barrier: they land while the CTAs wait
        st.rewind();
#pragma unroll
        for (int s = 0; s < S; ++s) st.issue(s);
        double* f_s = cta_stats;                                   // [K+1] (cta_stats is free until the tail)
        const PeerCtx ex1{a.ex_windows, a.ex_rank, a.ex_world, a.ex_max_len};
        // f_j and the label-change count are sums of values in [0, 1] over at most n points: fixed point with
        // 2^shift * n < 2^41, accumulated with integer atomics in the workspace header.  Integer addition is
        // associative, so the totals are bit-reproducible whatever the arrival order, and every WARP adds its own
        // sums (float -> warp shuffle tree -> fixed point): no CTA-level reduction in front of the grid barrier.
        // Every accumulator word also COUNTS its contributions (CountedFix, scc_common.cuh): a CTA knows the sums are
        // final by polling the K + 1 words themselves — no ticket, no separate read of the sums.
        unsigned long long* fix = reinterpret_cast<unsigned long long*>(a.counter) + kFixOffset;
        const int shift = CountedFix::kValueBits - 1 - (64 - __clzll((long long)(a.n > 0 ? a.n : 1)));
        const unsigned int seq1 = s_seq1;      // read in the prologue: no CTA can have advanced it by then
        {
            const double scale_fix = ldexp(1.0, shift);
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j <= KP; ++j) {
                const float v = warp_sum(j < KP ? ((j & 1) ? facc2[j / 2].y : facc2[j / 2].x) : changed);
                const int slot = (j < KP) ? j : K;
                if (lane == slot && (j == KP || j < K)) mine = v;
            }
            if (lane <= K) atomicAdd(fix + lane, CountedFix::word(mine, scale_fix, (double)a.n));
        }
        grid_barrier_counted<NT>(K + 1, fix, gridDim.x * NW, ldexp(1.0, -shift), f_s,
                                          reinterpret_cast<double*>(scratch), &ex1, seq1);
        if (threadIdx.x < KP) inv_f[threadIdx.x] = ((int)threadIdx.x < K) ? (float)(1.0 / f_s[threadIdx.x]) : 0.f;
        if (blockIdx.x == 0 && (int)threadIdx.x <= K && a.f_out) a.f_out[threadIdx.x] = f_s[threadIdx.x];
        __syncthreads();
        SCC_TL(a.timeline, 7);                                     // grid barrier + f all-reduce done
    }

    const float* krows = krow_operand<MODE>(a);
    float sm[2] = {0.f, 0.f};             // loss, sum s
    float2 B2[JP * DW];                   // (B_2jp,c , B_2jp+1,c) for c < D, then (W_2jp, W_2jp+1)
#pragma unroll
    for (int s = 0; s < JP * DW; ++s) B2[s] = make_float2(0.f, 0.f);
    {   // ---- gradient pass: no CTA-wide barrier in this loop
        int cons = 0, stage = 0, pending = -1;
        bool first = true;
        while (cons < st.total) {
            const int rows = st.stage_rows(cons);
            st.wait(stage, rows);
            if (first) { SCC_TL(a.timeline, 2); first = false; }
            float* sp = st.stage_ptr(stage);
            const int nsl = (rows + 31) >> 5;
            for (int r = 0; r < nsl; r += RB) {
                if (32 * r + lane < rows) {             // row k of the block: slice r + k (rows ascend: k = 0 is valid)
                    bool valid[RB];
                    int row_in[RB];
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
                        row_in[k] = 32 * (r + k) + lane;
                        valid[k] = row_in[k] < rows;
                    }
                    const size_t i0 = (size_t)st.row_begin + (cons + row_in[0]);
                    // the [n, K] operand rows (target p / upstream dL/dq) are requested before the z rows are read
                    float2 pre2[RB][JP];
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
#pragma unroll
                        for (int jp = 0; jp < JP; ++jp) pre2[k][jp] = make_float2(0.f, 0.f);
                        if (krows && valid[k]) load_krow2<KP, EXACT>(krows + (i0 + 32 * k) * K, K, pre2[k]);
                    }
                    float2 coef2[RB][JP];
                    {
                        float zr[RB][D];
#pragma unroll
                        for (int k = 0; k < RB; ++k) {
#pragma unroll
                            for (int c = 0; c < D; ++c) zr[k][c] = 0.f;
                            if (valid[k]) load_row<D>(sp, row_in[k], zr[k]);
                        }
                        float2 acc2[RB][JP];
                        sq_distances<D, KP, RB>(zr, nmuT2, acc2);
#pragma unroll
                        for (int k = 0; k < RB; ++k) {
#pragma unroll
                            for (int jp = 0; jp < JP; ++jp) coef2[k][jp] = make_float2(0.f, 0.f);
                            if (valid[k]) {
                                float2 w2[JP], u2[JP], t2[JP];
                                int label;
                                float best, tsum;
                                student_t_pairs<KP, EXACT, ALPHA1, MODE == MODE_KMEANS>(acc2[k], K, inv_alpha, expo, w2, u2, t2,
                                                                                        tsum, label, best);
                                grad_coefficients<KP, EXACT, ALPHA1, MODE>(a, i0 + 32 * k, K, inv_f2, w2, u2, t2, tsum, expo,
                                                                           label, best, pre2[k], coef2[k], sm[0], sm[1]);
                            }
                        }
                    }
                    // the rows are read a second time for the accumulation phase: keeping them in registers across the
                    // coefficient phase costs D registers per row the budget does not have (spills)
                    asm volatile("" ::: "memory");
                    float2 zc2[RB][DP2];                           // centred points as dimension pairs
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
                        float zb[D];
#pragma unroll
                        for (int c = 0; c < D; ++c) zb[c] = 0.f;
                        if (valid[k]) load_row<D>(sp, row_in[k], zb);
#pragma unroll
                        for (int c = 0; c < DP2; ++c)
                            zc2[k][c] = __fadd2_rn(make_float2(zb[2 * c], (2 * c + 1 < D) ? zb[2 * c + 1] : 0.f), nc0_s[c]);
                    }
#pragma unroll
                    for (int k = 0; k < RB; ++k) {                 // row after row: the summation order of RB = 1
#pragma unroll
                        for (int c = 0; c < D; ++c) {
                            const float zc = (c & 1) ? zc2[k][c >> 1].y : zc2[k][c >> 1].x;
#pragma unroll
                            for (int jp = 0; jp < JP; ++jp)
                                B2[jp * DW + c] = __ffma2_rn(coef2[k][jp], make_float2(zc, zc), B2[jp * DW + c]);
                        }
#pragma unroll
                        for (int jp = 0; jp < JP; ++jp) B2[jp * DW + D] = __fadd2_rn(B2[jp * DW + D], coef2[k][jp]);
                    }
                    if (want_dz) {
                        float csum[RB];
#pragma unroll
                        for (int k = 0; k < RB; ++k) {
                            float2 c2 = make_float2(0.f, 0.f);
#pragma unroll
                            for (int jp = 0; jp < JP; ++jp) c2 = __fadd2_rn(c2, coef2[k][jp]);
                            csum[k] = cs * (c2.x + c2.y);
                        }
                        float dzr[RB][D];
                        dz_from_coefficients_rows<D, KP, EXACT, RB>(zc2, coef2, csum, nmc2_s, K, dzr);
#pragma unroll
                        for (int k = 0; k < RB; ++k)
                            if (valid[k]) store_row<D>(sp, row_in[k], dzr[k]);     // in place: the z rows are in registers
                    }
                }
                if (r == 0 && pending >= 0) {           // refill the stage consumed before this one: its dz store
                    if (kBulkOut && want_dz && lane == 0) bulk_wait_read0();     // has read shared memory by now
                    __syncwarp();
                    st.issue(pending);
                    pending = -1;
                }
            }
            if (want_dz) {
                float* dst = a.dz + ((size_t)st.row_begin + cons) * D;
                if (kBulkOut && Stream::tma_ok(rows)) {
                    fence_proxy_async();                // the warp's dz rows -> async proxy -> one bulk store by lane 0
                    __syncwarp();
                    if (lane == 0) {
                        bulk_s2g(dst, sp, (uint32_t)rows * D * sizeof(float));
                        bulk_commit();
                    }
                } else {
                    __syncwarp();
                    warp_copy_rows_out<D>(sp, dst, rows);
                }
            }
            __syncwarp();
            cons += rows;
            pending = stage;
            if (++stage == S) stage = 0;
        }
    }
    pdl_trigger();                      // successor may start its prologue under our reduction tail
    SCC_TL(a.timeline, 3);
    if (mode_is_kl<MODE>()) sm[0] *= a.scale * 0.693147180559945f;     // loss = scale * ln2 * sum p log2(p/q)
    float acc[NV];
    acc[0] = sm[0]; acc[1] = sm[1];
#pragma unroll
    for (int j = 0; j < KP; ++j) acc[2 + j] = (j & 1) ? B2[(j / 2) * DW + D].y : B2[(j / 2) * DW + D].x;
#pragma unroll
    for (int j = 0; j < KP; ++j)
#pragma unroll
        for (int c = 0; c < D; ++c)
            acc[2 + KP + j * D + c] = (j & 1) ? B2[(j / 2) * DW + c].y : B2[(j / 2) * DW + c].x;
    __syncthreads();
    cta_reduce<NV, NT>(acc, scratch, cta_stats);
    SCC_TL(a.timeline, 4);
    // dmu_jc = -cs (B_jc - W_j (mu_jc - c0_c)), compacted to [loss, sum s, dmu[K*D]]
    // (MODE_KMEANS: [inertia, 0, sum_{i in j} (z_i - mu_j) [K*D], counts[K]])
    static_assert(KP * D <= NT, "one thread per statistic in the tail");
    double dmu = 0.0, wj = 0.0;
    const int o = threadIdx.x;
    if (o < K * D) dmu = -(double)cs * (cta_stats[2 + KP + o] - cta_stats[2 + o / D] * (double)mc_s[o]);
    if (o < K) wj = cta_stats[2 + o];
    __syncthreads();
    if (o < K * D) cta_stats[2 + o] = (MODE == MODE_KMEANS) ? -dmu : dmu;
    if (MODE == MODE_KMEANS && o < K) cta_stats[2 + K * D + o] = wj;
    __syncthreads();
    const PeerCtx push{a.ex_push ? a.ex_windows : nullptr, a.ex_rank, a.ex_world, a.ex_max_len};
    // (the warp's last dz bulk stores read the stream buffers, which nothing in the tail touches: their completion
    //  is only awaited at the very end, behind the reductions)
    const bool last = grid_publish<NT, 30>(cta_stats, K * D + 2 + (MODE == MODE_KMEANS ? K : 0), a.partials,
                                                    a.counter, a.stats, scratch, &push, a.ex_push);
    if (MODE == MODE_STEP && last) {                                          // every CTA is past the pass-1 barrier
        if ((int)threadIdx.x <= K) reinterpret_cast<unsigned long long*>(a.counter)[kFixOffset + threadIdx.x] = 0ull;
    }
    if (kBulkOut && lane == 0) bulk_wait0();             // this warp's dz stores are complete
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
    // output blocks of W^T Z per lane: 8x8 where the shape allows (two LDS.128 of W + two of Z per 64 FMAs), else 4x4
    constexpr int BR = (KP % 8 == 0 && D % 8 == 0) ? 8 : 4;     // clusters per block
    constexpr int BC = BR;                                      // dimensions per block
    constexpr int NB = (KP / BR) * (D / BC);
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
    constexpr int JP = KP / 2;
    float2* nmuT2 = reinterpret_cast<float2*>(w_tile + kDecTile * KP);       // [D][JP] (D even: same size as [KP][DP2])
    float2* nmc2_s = nmuT2 + D * JP;                         // [KP][DP2]
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
    load_grad_constants<D, KP>(a, K, cs, nmuT2, nmc2_s, mc_s, c0_s, inv_f);
    __syncthreads();

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    const bool want_dz = a.dz != nullptr;
    const float2* inv_f2 = reinterpret_cast<const float2*>(inv_f);
    float2 nc0[DP2];
#pragma unroll
    for (int c = 0; c < DP2; ++c) nc0[c] = make_float2(-c0_s[2 * c], -c0_s[2 * c + 1]);
    float small[NSM];
#pragma unroll
    for (int s = 0; s < NSM; ++s) small[s] = 0.f;
    float blk[BR * BC];
#pragma unroll
    for (int s = 0; s < BR * BC; ++s) blk[s] = 0.f;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / NB, lb = lane - grp * NB;  // lanes >= G2*NB idle in phase 2
    const int jb = lb / (D / BC), cb = lb - jb * (D / BC);
    const bool p2_active = grp < G2;

    const float* krows = MODE == MODE_KLU ? a.u_in : krow_operand<MODE>(a);
    float2 nxt2[JP];                     // operand row of the next tile (requested before phase 2: latency hidden)
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) nxt2[jp] = make_float2(0.f, 0.f);
    if (krows && (int)blockIdx.x < ring.num_tiles && (int)threadIdx.x < ring.points(blockIdx.x))
        load_krow2<KP, EXACT>(krows + ((size_t)blockIdx.x * kDecTile + threadIdx.x) * K, K, nxt2);
    int stage = 0;
    uint32_t use = 0;
    for (int tile = blockIdx.x; tile < ring.num_tiles; tile += G) {
        const int np = ring.points(tile);
        const int64_t base = (int64_t)tile * kDecTile;
        const bool active = (int)threadIdx.x < np;
        // the [n, K] operand row (target p / upstream dL/dq / handed-over u) was requested one tile ahead
        float2 pre2[JP];
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) pre2[jp] = nxt2[jp];
        ring.wait(stage, tile, use);
        float* ztile = ring.stage_ptr(stage);
        float2 coef2[JP];
#pragma unroll
        for (int jp = 0; jp < JP; ++jp) coef2[jp] = make_float2(0.f, 0.f);
        if (active) {
            float zr[1][D];
            load_row<D>(ztile, threadIdx.x, zr[0]);
            const size_t i = (size_t)base + threadIdx.x;
            float2 w2[JP], u2[JP], t2[JP];
            int label = 0;
            float best = 0.f, tsum;
            if constexpr (MODE == MODE_KLU) {
                // u_ij = 1 / (1 + d_ij / alpha) handed over by the assign pass: no distance loop (the centroid table
                // reads of that loop are what bounds this kernel: shared-memory return bandwidth)
#pragma unroll
                for (int jp = 0; jp < JP; ++jp) u2[jp] = pre2[jp];
                float2 ts2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int jp = 0; jp < JP; ++jp) {
                    const bool vx = EXACT || 2 * jp < K, vy = EXACT || 2 * jp + 1 < K;
                    const float2 uu = u2[jp];
                    float2 ww = make_float2(rcp_approx(vx ? uu.x : 1.f), rcp_approx(vy ? uu.y : 1.f));
                    float2 tt = ALPHA1 ? uu : make_float2(ex2_approx(expo * lg2_approx(vx ? uu.x : 1.f)),
                                                          ex2_approx(expo * lg2_approx(vy ? uu.y : 1.f)));
                    if (!EXACT) { tt = pair_sel(vx, vy, tt, 0.f); u2[jp] = pair_sel(vx, vy, uu, 0.f); }
                    w2[jp] = ww; t2[jp] = tt;
                    ts2 = __fadd2_rn(ts2, tt);
                }
                tsum = ts2.x + ts2.y;
            } else {
                float2 acc2[1][JP];
                sq_distances<D, KP, 1>(zr, nmuT2, acc2);
                student_t_pairs<KP, EXACT, ALPHA1, MODE == MODE_KMEANS>(acc2[0], K, inv_alpha, expo, w2, u2, t2, tsum,
                                                                        label, best);
            }
            grad_coefficients<KP, EXACT, ALPHA1, MODE>(a, i, K, inv_f2, w2, u2, t2, tsum, expo, label, best, pre2, coef2,
                                                       small[0], small[1]);
            float2 c2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int jp = 0; jp < JP; ++jp) {
                small[2 + 2 * jp] += coef2[jp].x; small[3 + 2 * jp] += coef2[jp].y;
                c2 = __fadd2_rn(c2, coef2[jp]);
            }
            float2 z2[DP2];
#pragma unroll
            for (int c = 0; c < DP2; ++c) z2[c] = __fadd2_rn(make_float2(zr[0][2 * c], zr[0][2 * c + 1]), nc0[c]);
#pragma unroll
            for (int c = 0; c < D; ++c) zr[0][c] = (c & 1) ? z2[c >> 1].y : z2[c >> 1].x;
            store_row<D>(ztile, threadIdx.x, zr[0]);  // own row, centred, for phase 2
            if (want_dz) {
                float dzr[D];
                dz_from_coefficients<D, KP, EXACT>(z2, coef2, cs * (c2.x + c2.y), nmc2_s, K, dzr);
                store_row<D>(out_tile, threadIdx.x, dzr);
            }
        }
#pragma unroll
        for (int j = 0; j < KP; j += 4)
            *reinterpret_cast<float4*>(w_tile + threadIdx.x * KP + j) =
                make_float4(coef2[j / 2].x, coef2[j / 2].y, coef2[j / 2 + 1].x, coef2[j / 2 + 1].y);
        __syncthreads();
        {   // request the next tile's operand row now: it lands during phase 2
            const int nt = tile + G;
            if (krows && nt < ring.num_tiles && (int)threadIdx.x < ring.points(nt))
                load_krow2<KP, EXACT>(krows + ((size_t)nt * kDecTile + threadIdx.x) * K, K, nxt2);
        }
        if (want_dz) copy_tile_out<D, kDecThreads>(out_tile, a.dz + (size_t)base * D, np);
        if (p2_active) {
            for (int r = warp * G2 + grp; r < np; r += NW * G2) {
                // this phase is bound by its shared-memory reads (return bandwidth 128 B/clk/SM): BR/4 + BC/4 LDS.128
                // per BR*BC FMAs.  Scalar FMAs on purpose: the packed form was measured slower here (round 1).
                float wv[BR], xv[BC];
#pragma unroll
                for (int h = 0; h < BR / 4; ++h) {
                    const float4 w = *reinterpret_cast<const float4*>(w_tile + r * KP + BR * jb + 4 * h);
                    wv[4 * h] = w.x; wv[4 * h + 1] = w.y; wv[4 * h + 2] = w.z; wv[4 * h + 3] = w.w;
                }
#pragma unroll
                for (int h = 0; h < BC / 4; ++h) {
                    const float4 x = *reinterpret_cast<const float4*>(ztile + r * L::LD + BC * cb + 4 * h);
                    xv[4 * h] = x.x; xv[4 * h + 1] = x.y; xv[4 * h + 2] = x.z; xv[4 * h + 3] = x.w;
                }
#pragma unroll
                for (int rr = 0; rr < BR; ++rr)
#pragma unroll
                    for (int cc = 0; cc < BC; ++cc) blk[rr * BC + cc] = fmaf(wv[rr], xv[cc], blk[rr * BC + cc]);
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
    // per-(warp, group) block partials -> shared (the ring buffer is free now), fixed-order sum in float64
    float* part = ring_buf;                                                  // [NW*G2][KP*D] (fp32 values as accumulated)
    if (p2_active) {
#pragma unroll
        for (int r = 0; r < BR; ++r)
#pragma unroll
            for (int c = 0; c < BC; ++c)
                part[(size_t)(warp * G2 + grp) * (KP * D) + (BR * jb + r) * D + BC * cb + c] = blk[r * BC + c];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < K * D; o += kDecThreads) {
        double accd = 0.0;
#pragma unroll
        for (int g = 0; g < NW * G2; ++g) accd += (double)part[(size_t)g * (KP * D) + o];
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
    return sizeof(float) * (((S * kDecTile * assign_ppt<D, KP>() * RowLayout<D>::LD + 3) & ~3) + 2 * ((D * (KP / 2) + 1) & ~1)) +
           sizeof(double) * (KP + 1) + sizeof(uint64_t) * S;
}
template <int D, int KP>
constexpr size_t grad_tiled_smem() {
    constexpr int S = 2;
    constexpr int BR = (KP % 8 == 0 && D % 8 == 0) ? 8 : 4;
    constexpr int NB = (KP / BR) * (D / BR);
    constexpr int G2 = (32 / NB) > 0 ? (32 / NB) : 1;
    constexpr int NW = kDecThreads / 32;
    constexpr int NSM = KP + 2;
    constexpr int scr = reduce_scratch(NSM);
    size_t bytes = sizeof(float) * ((S + 1) * kDecTile * RowLayout<D>::LD + kDecTile * KP + 4 * KP * Pairs<D>::N +
                                    KP * D + D + KP) +
                   sizeof(double) * (scr + NSM + KP * D + 2 + KP) + sizeof(uint64_t) * S;
    // the ring buffer is reused for the [NW*G2][KP*D] float partials at the end
    const size_t part = sizeof(float) * NW * G2 * KP * D;
    const size_t ring = sizeof(float) * S * kDecTile * RowLayout<D>::LD;
    if (part > ring) bytes += part - ring;
    return bytes;
}

template <typename Kern>
static int launch_dec(Kern kern, const DecArgs& args, size_t smem, cudaStream_t stream, int tile_points = kDecTile,
                      bool cooperative = false, int threads = kDecThreads) {
    const int64_t num_tiles = (args.n + tile_points - 1) / tile_points;
    int64_t grid = persistent_grid(reinterpret_cast<const void*>(kern), threads, smem, kMaxCtasPerSm);
    if (grid < 0) return (int)grid;
    if (grid > kMaxDecGrid) grid = kMaxDecGrid;
    if (grid > num_tiles) grid = num_tiles;
    if (args.batch > 0 && grid > kBatchGridX) grid = kBatchGridX;    // R restarts share the machine (workspace bound)
    if (cooperative && grid * (threads / 32) > (int64_t)CountedFix::kMaxContributors)
        grid = CountedFix::kMaxContributors / (threads / 32);      // one counted contribution per warp at the grid barrier
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid, args.batch > 0 ? (unsigned)args.batch : 1u);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int nattr = 1;
    if (cooperative) {              // grid-wide barrier inside: every CTA must be resident (grid <= SMs x occupancy)
        // Cooperative launch (co-residency guaranteed by the driver) + programmatic dependent launch: the next
        // kernel's CTAs are scheduled while this kernel's reduction tail drains and wait in griddepcontrol.wait
        // (step 47.3 -> 46.4 us in 20-step graphs).  SCC_STEP_LAUNCH=0 drops the PDL attribute; 2 (experiments
        // only) drops the cooperative one.
        static const int step_launch = [] {
            const char* e = getenv("SCC_STEP_LAUNCH");
            return e ? atoi(e) : 1;
        }();
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        if (step_launch >= 1) {
            attr[step_launch == 1 ? 1 : 0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[step_launch == 1 ? 1 : 0].val.programmaticStreamSerializationAllowed = 1;
            nattr = step_launch == 1 ? 2 : 1;
        }
    } else {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // PDL, see scc_common.cuh
        attr[0].val.programmaticStreamSerializationAllowed = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = nattr;
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
            if constexpr (MODE == MODE_KLU)
                return SCC_ERR_UNSUPPORTED;          // register-blocked shapes run the one-kernel step instead
            else
                return launch_dec(dec_grad_reg_kernel<D, KP, EXACT, A1, MODE>, a, grad_reg_smem<D, KP>(), st,
                                  reg_threads<D, KP>(), MODE == MODE_STEP, reg_threads<D, KP>());
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
