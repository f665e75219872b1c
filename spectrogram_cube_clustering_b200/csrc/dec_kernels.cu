// dec_kernels.cu — DEC clustering-layer kernels for sm_100a (CUDA cores, HBM-bound).
//
//   dec_assign_kernel   z -> q, labels, f_j = sum_i q_ij, label-change count   (one read of z)
//                       replaces Cluster/networks.py:279-288 + models.py:92,94,1098-1099,1320
//   dec_target_kernel   q, f -> p                                               (models.py:1320-1322)
//   dec_grad_kernel     z (+p | +f | +dL/dq) -> loss, dz, dmu                   (models.py:1124-1127 + autograd)
//       REG   variant: per-thread register accumulators for dmu   (K*d <= 160)
//       TILED variant: warp-level 4x4 register-blocked W^T Z over the staged tile (d % 4 == 0)
//
// One thread owns one latent point: its row sits in registers, centroids are
// broadcast from shared memory, q / coefficients never leave registers.
#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

constexpr int kDecThreads = 256;
constexpr int kDecTile = 256;

struct DecArgs {
    const float* z;
    int64_t n;
    const float* mu;
    int K;
    float alpha;
    int round5;
    // assign
    float* q;
    int32_t* labels;
    const int32_t* labels_prev;
    // grad
    const float* p;
    const double* f_cols;
    const float* grad_q;
    float scale;
    float* dz;
    // reduction
    double* stats;
    double* partials;
    unsigned int* counter;
};

template <int D>
__host__ __device__ constexpr int dec_stages() { return RowLayout<D>::kDense ? 4 : 2; }

// ---------------------------------------------------------------------------
// q_i, u_i and the hard label of one point.  networks.py:279-288, models.py:92.
// ---------------------------------------------------------------------------
template <int D, int KP, bool ALPHA1>
__device__ __forceinline__ void soft_assign_row(const float (&zr)[D], const float* __restrict__ mu_s, int K,
                                                float inv_alpha, float expo, float (&u)[KP], float (&q)[KP],
                                                int& label) {
    float tsum = 0.f, best = 3.4e38f;
    label = 0;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        u[j] = 0.f; q[j] = 0.f;
        if (j < K) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float df = zr[c] - mu_s[j * D + c];
                acc = fmaf(df, df, acc);
            }
            if (acc < best) { best = acc; label = j; }          // argmax q == argmin distance, first wins
            const float uu = __fdividef(1.f, fmaf(acc, inv_alpha, 1.f));
            const float t = ALPHA1 ? uu : __powf(uu, expo);
            u[j] = uu; q[j] = t; tsum += t;
        }
    }
    const float inv = __fdividef(1.f, tsum);
#pragma unroll
    for (int j = 0; j < KP; ++j) q[j] *= inv;
}

template <int KP>
__device__ __forceinline__ void store_krow(float* __restrict__ dst, int K, const float (&v)[KP]) {
    if ((K & 3) == 0) {
#pragma unroll
        for (int j = 0; j < KP; j += 4)
            if (j < K) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < KP; ++j)
            if (j < K) dst[j] = v[j];
    }
}
template <int KP>
__device__ __forceinline__ void load_krow(const float* __restrict__ src, int K, float (&v)[KP]) {
    if ((K & 3) == 0) {
#pragma unroll
        for (int j = 0; j < KP; j += 4) {
            if (j < K) {
                const float4 x = ldg_stream4(reinterpret_cast<const float4*>(src + j));
                v[j] = x.x; v[j + 1] = x.y; v[j + 2] = x.z; v[j + 3] = x.w;
            } else { v[j] = 0.f; if (j + 1 < KP) v[j + 1] = 0.f; if (j + 2 < KP) v[j + 2] = 0.f; if (j + 3 < KP) v[j + 3] = 0.f; }
        }
    } else {
#pragma unroll
        for (int j = 0; j < KP; ++j) v[j] = (j < K) ? ldg_stream(src + j) : 0.f;
    }
}

// Reduce NV per-thread floats across the CTA into cta_stats[base .. base+NV) (float64).
// scratch: [num_warps][NV] doubles.  Deterministic (fixed warp order).
template <int NV, int NT>
__device__ __forceinline__ void cta_reduce(const float (&v)[NV], double* scratch, double* cta_stats, int base) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 0; s < NV; ++s) {
        const float w = warp_sum(v[s]);
        if (lane == 0) scratch[warp * NV + s] = (double)w;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < NV; s += NT) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) acc += scratch[w * NV + s];
        cta_stats[base + s] = acc;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// dec_assign
// ---------------------------------------------------------------------------
template <int D, int KP, bool ALPHA1>
__global__ void __launch_bounds__(kDecThreads)
dec_assign_kernel(const DecArgs a) {
    constexpr int S = dec_stages<D>();
    using Ring = ZRing<D, kDecTile, S, kDecThreads>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* mu_s = ring_buf + S * Ring::kTileFloats;                         // [KP*D]
    double* scratch = reinterpret_cast<double*>(mu_s + ((KP * D + 3) & ~3)); // [8][KP+1]
    double* cta_stats = scratch + (kDecThreads / 32) * (KP + 1);            // [KP+1]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + (KP + 1));

    const int K = a.K;
    for (int i = threadIdx.x; i < KP * D; i += kDecThreads) mu_s[i] = (i < K * D) ? a.mu[i] : 0.f;

    Ring ring;
    ring.init(ring_buf, bars, a.z, a.n);
    __syncthreads();
    for (int s = 0; s < S; ++s) ring.issue(s, (int64_t)blockIdx.x + (int64_t)s * gridDim.x);
    __syncthreads();

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    float facc[KP + 1];
#pragma unroll
    for (int j = 0; j <= KP; ++j) facc[j] = 0.f;

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ring.num_tiles; tile += gridDim.x, ++it) {
        const int stage = it % S;
        ring.wait(stage, tile, (uint32_t)(it / S));
        const int np = ring.points(tile);
        const bool active = (int)threadIdx.x < np;
        float zr[D];
        if (active) load_row<D>(ring.stage_ptr(stage), threadIdx.x, zr);
        __syncthreads();
        ring.issue(stage, tile + (int64_t)S * gridDim.x);
        if (active) {
            const int64_t i = tile * kDecTile + threadIdx.x;
            float u[KP], q[KP];
            int label;
            soft_assign_row<D, KP, ALPHA1>(zr, mu_s, K, inv_alpha, expo, u, q, label);
            if (a.round5) {
#pragma unroll
                for (int j = 0; j < KP; ++j) q[j] = round_dec5(q[j]);
            }
#pragma unroll
            for (int j = 0; j < KP; ++j) facc[j] += q[j];
            if (a.q) store_krow<KP>(a.q + i * K, K, q);
            if (a.labels) a.labels[i] = label;
            if (a.labels_prev) facc[KP] += (a.labels_prev[i] != label) ? 1.f : 0.f;
        }
    }
    cta_reduce<KP + 1, kDecThreads>(facc, scratch, cta_stats, 0);
    // stats layout is [K+1]: compact the KP-padded vector
    if (threadIdx.x == 0 && K < KP) cta_stats[K] = cta_stats[KP];
    __syncthreads();
    grid_publish(cta_stats, K + 1, a.partials, a.counter, a.stats);
}

// ---------------------------------------------------------------------------
// dec_target: p = normalise_rows(q^2 / f)   (models.py:1320-1322)
// LPR lanes cooperate on one row (K = 4*LPR) so that global accesses are
// 128-bit and fully coalesced; LPR = 0 is the scalar thread-per-row fallback.
// ---------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256)
dec_target_kernel(const float* __restrict__ q, int64_t n, int K, const double* __restrict__ f,
                  int round5, float* __restrict__ p) {
    __shared__ float inv_f[SCC_MAX_K];
    if (threadIdx.x < K) inv_f[threadIdx.x] = (float)(1.0 / f[threadIdx.x]);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if constexpr (LPR > 0) {
        const int64_t nvec = n * LPR;
        const int sub = threadIdx.x % LPR;       // blockDim (256) is a multiple of LPR, so is the grid stride
        const float i0 = inv_f[4 * sub], i1 = inv_f[4 * sub + 1], i2 = inv_f[4 * sub + 2], i3 = inv_f[4 * sub + 3];
        const int64_t nvec_pad = (nvec + 31) & ~int64_t(31);
        for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec_pad; v += stride) {
            const bool ok = v < nvec;
            float4 x = ok ? ldg_stream4(reinterpret_cast<const float4*>(q) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 w = make_float4(x.x * x.x * i0, x.y * x.y * i1, x.z * x.z * i2, x.w * x.w * i3);
            float s = (w.x + w.y) + (w.z + w.w);
#pragma unroll
            for (int o = 1; o < LPR; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float inv = 1.f / s;
            w.x *= inv; w.y *= inv; w.z *= inv; w.w *= inv;
            if (round5) { w.x = round_dec5(w.x); w.y = round_dec5(w.y); w.z = round_dec5(w.z); w.w = round_dec5(w.w); }
            if (ok) reinterpret_cast<float4*>(p)[v] = w;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            float w[SCC_MAX_K];
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < SCC_MAX_K; ++j) {
                if (j < K) { const float x = q[i * K + j]; w[j] = x * x * inv_f[j]; s += w[j]; }
            }
            const float inv = 1.f / s;
#pragma unroll
            for (int j = 0; j < SCC_MAX_K; ++j) {
                if (j < K) { float v = w[j] * inv; if (round5) v = round_dec5(v); p[i * K + j] = v; }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Per-point gradient coefficients c_ij with dz_i = sum_j c_ij (z_i - mu_j),
// dmu_j = -sum_i c_ij (z_i - mu_j).
//   MODE_KL      : c_ij = scale (alpha+1)/alpha (p_ij - q_ij s_i) u_ij   (+ loss)
//   MODE_GENERIC : c_ij = -(alpha+1)/alpha q_ij (G_ij - sum_j G_ij q_ij) u_ij
// ---------------------------------------------------------------------------
enum { MODE_KL = 0, MODE_GENERIC = 1 };

template <int KP, int MODE>
__device__ __forceinline__ void grad_coefficients(const DecArgs& a, int64_t i, int K, const float* __restrict__ inv_f,
                                                  const float (&u)[KP], const float (&q)[KP], float cscale,
                                                  float (&coef)[KP], float& loss, float& ssum) {
    if constexpr (MODE == MODE_KL) {
        float p[KP];
        if (a.p) {
            load_krow<KP>(a.p + i * K, K, p);
        } else {                                   // rebuild p from the column sums (fused mode)
            float wsum = 0.f;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const float qq = a.round5 ? round_dec5(q[j]) : q[j];
                p[j] = (j < K) ? qq * qq * inv_f[j] : 0.f;
                wsum += p[j];
            }
            const float inv = 1.f / wsum;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                p[j] *= inv;
                if (a.round5) p[j] = round_dec5(p[j]);
            }
        }
        float s = 0.f, l = 0.f;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            if (j < K) {
                s += p[j];
                // xlogy(p,p) - p log q ; 0 when p == 0 (torch KLDivLoss)
                const float term = p[j] * (__log2f(__fdividef(p[j], q[j])) * 0.693147180559945f);
                l += (p[j] == 0.f) ? 0.f : term;            // NaN targets still propagate
            }
        }
#pragma unroll
        for (int j = 0; j < KP; ++j) coef[j] = (j < K) ? (p[j] - q[j] * s) * u[j] * cscale : 0.f;
        loss += l; ssum += s;
    } else {
        float g[KP];
        load_krow<KP>(a.grad_q + i * K, K, g);
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < KP; ++j) dot = fmaf(g[j], q[j], dot);
#pragma unroll
        for (int j = 0; j < KP; ++j) coef[j] = (j < K) ? -cscale * q[j] * (g[j] - dot) * u[j] : 0.f;
    }
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

// ---------------------------------------------------------------------------
// dec_grad, REG variant: dmu accumulated in per-thread registers.
// stats out: [loss, sum_i s_i, dmu[K*D]]
// ---------------------------------------------------------------------------
template <int D, int KP, bool ALPHA1, int MODE>
__global__ void __launch_bounds__(kDecThreads)
dec_grad_reg_kernel(const DecArgs a) {
    constexpr int S = dec_stages<D>();
    using Ring = ZRing<D, kDecTile, S, kDecThreads>;
    using L = RowLayout<D>;
    constexpr int NV = KP * D + 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* out_tile = ring_buf + S * Ring::kTileFloats;                      // [TILE*LD]
    float* mu_s = out_tile + Ring::kTileFloats;                              // [KP*D]
    float* inv_f = mu_s + ((KP * D + 3) & ~3);                               // [KP]
    double* scratch = reinterpret_cast<double*>(inv_f + ((KP + 3) & ~3));    // [8][NV]
    double* cta_stats = scratch + (kDecThreads / 32) * NV;                   // [NV]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + NV);

    const int K = a.K;
    for (int i = threadIdx.x; i < KP * D; i += kDecThreads) mu_s[i] = (i < K * D) ? a.mu[i] : 0.f;
    if (threadIdx.x < KP)
        inv_f[threadIdx.x] = (a.f_cols && (int)threadIdx.x < K) ? (float)(1.0 / a.f_cols[threadIdx.x]) : 0.f;

    Ring ring;
    ring.init(ring_buf, bars, a.z, a.n);
    __syncthreads();
    for (int s = 0; s < S; ++s) ring.issue(s, (int64_t)blockIdx.x + (int64_t)s * gridDim.x);
    __syncthreads();

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    const float cscale = (MODE == MODE_KL ? a.scale : 1.f) * (a.alpha + 1.f) / a.alpha;
    float acc[NV];                                   // [0]=loss [1]=sum s [2..]=sum_i c_ij (z_i - mu_j)
#pragma unroll
    for (int s = 0; s < NV; ++s) acc[s] = 0.f;

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ring.num_tiles; tile += gridDim.x, ++it) {
        const int stage = it % S;
        ring.wait(stage, tile, (uint32_t)(it / S));
        const int np = ring.points(tile);
        const bool active = (int)threadIdx.x < np;
        float zr[D];
        if (active) load_row<D>(ring.stage_ptr(stage), threadIdx.x, zr);
        __syncthreads();                 // stage free; previous out_tile fully copied out
        ring.issue(stage, tile + (int64_t)S * gridDim.x);
        if (active) {
            const int64_t i = tile * kDecTile + threadIdx.x;
            float u[KP], q[KP], coef[KP];
            int label;
            soft_assign_row<D, KP, ALPHA1>(zr, mu_s, K, inv_alpha, expo, u, q, label);
            grad_coefficients<KP, MODE>(a, i, K, inv_f, u, q, cscale, coef, acc[0], acc[1]);
            float dzr[D];
#pragma unroll
            for (int c = 0; c < D; ++c) dzr[c] = 0.f;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                if (j < K) {
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        const float df = zr[c] - mu_s[j * D + c];
                        dzr[c] = fmaf(coef[j], df, dzr[c]);
                        acc[2 + j * D + c] = fmaf(coef[j], df, acc[2 + j * D + c]);
                    }
                }
            }
            if (a.dz) store_row<D>(out_tile, threadIdx.x, dzr);
        }
        if (a.dz) {
            __syncthreads();
            copy_tile_out<D, kDecThreads>(out_tile, a.dz + tile * (int64_t)kDecTile * D, np);
        }
    }
    // dmu_j = -sum_i c_ij (z_i - mu_j);  loss = scale * sum p log(p/q)
    acc[0] *= a.scale;
#pragma unroll
    for (int s = 2; s < NV; ++s) acc[s] = -acc[s];
    __syncthreads();
    cta_reduce<NV, kDecThreads>(acc, scratch, cta_stats, 0);
    grid_publish(cta_stats, K * D + 2, a.partials, a.counter, a.stats);
}

// ---------------------------------------------------------------------------
// dec_grad, TILED variant (D % 4 == 0, KP % 4 == 0): phase 1 = thread per point
// (coefficients + dz), phase 2 = each warp accumulates W^T (Z - c0) over its
// share of the tile with 4x4 register blocks (one LDS.128 of W and one of Z per
// 16 FMAs).  c0 = mean centroid, removed to keep the sums well conditioned:
//   dmu_jc = -( A_jc - Wsum_j (mu_jc - c0_c) ),  A = sum_i c_ij (z_ic - c0_c).
// ---------------------------------------------------------------------------
template <int D, int KP, bool ALPHA1, int MODE>
__global__ void __launch_bounds__(kDecThreads)
dec_grad_tiled_kernel(const DecArgs a) {
    static_assert(D % 4 == 0 && KP % 4 == 0, "tiled variant needs 4-aligned shapes");
    constexpr int S = 2;
    using Ring = ZRing<D, kDecTile, S, kDecThreads>;
    using L = RowLayout<D>;
    constexpr int NB = (KP / 4) * (D / 4);            // 4x4 output blocks
    constexpr int G = (32 / NB) > 0 ? (32 / NB) : 1;  // point groups per warp
    constexpr int NW = kDecThreads / 32;
    constexpr int NS = KP * D + 2;
    static_assert(NB <= 32, "too many output blocks for one warp");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring_buf = reinterpret_cast<float*>(smem_raw);
    float* out_tile = ring_buf + S * Ring::kTileFloats;      // [TILE*LD] dz staging
    float* w_tile = out_tile + Ring::kTileFloats;            // [TILE*KP] coefficients
    float* mu_s = w_tile + kDecTile * KP;                    // [KP*D]
    float* c0_s = mu_s + KP * D;                             // [D]
    float* inv_f = c0_s + D;                                 // [KP]
    double* scratch = reinterpret_cast<double*>(inv_f + KP); // [NW][KP+2]
    double* cta_stats = scratch + NW * (KP + 2);             // [NS + KP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(cta_stats + NS + KP + 2);

    const int K = a.K;
    for (int i = threadIdx.x; i < KP * D; i += kDecThreads) mu_s[i] = (i < K * D) ? a.mu[i] : 0.f;
    if (threadIdx.x < KP)
        inv_f[threadIdx.x] = (a.f_cols && (int)threadIdx.x < K) ? (float)(1.0 / a.f_cols[threadIdx.x]) : 0.f;
    if (threadIdx.x < D) {
        float m = 0.f;
        for (int j = 0; j < K; ++j) m += a.mu[j * D + threadIdx.x];
        c0_s[threadIdx.x] = m / (float)K;
    }

    Ring ring;
    ring.init(ring_buf, bars, a.z, a.n);
    __syncthreads();
    for (int s = 0; s < S; ++s) ring.issue(s, (int64_t)blockIdx.x + (int64_t)s * gridDim.x);
    __syncthreads();

    const float inv_alpha = 1.f / a.alpha, expo = 0.5f * (a.alpha + 1.f);
    const float cscale = (MODE == MODE_KL ? a.scale : 1.f) * (a.alpha + 1.f) / a.alpha;
    float small[KP + 2];                              // loss, sum s, Wsum_j
#pragma unroll
    for (int s = 0; s < KP + 2; ++s) small[s] = 0.f;
    float blk[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) blk[s] = 0.f;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / NB, lb = lane - grp * NB;  // lanes >= G*NB idle in phase 2
    const int jb = lb / (D / 4), cb = lb - jb * (D / 4);
    const bool p2_active = grp < G;

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ring.num_tiles; tile += gridDim.x, ++it) {
        const int stage = it % S;
        ring.wait(stage, tile, (uint32_t)(it / S));
        const int np = ring.points(tile);
        const bool active = (int)threadIdx.x < np;
        float* ztile = ring.stage_ptr(stage);
        float coef[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) coef[j] = 0.f;
        if (active) {
            float zr[D];
            load_row<D>(ztile, threadIdx.x, zr);
            const int64_t i = tile * kDecTile + threadIdx.x;
            float u[KP], q[KP];
            int label;
            soft_assign_row<D, KP, ALPHA1>(zr, mu_s, K, inv_alpha, expo, u, q, label);
            grad_coefficients<KP, MODE>(a, i, K, inv_f, u, q, cscale, coef, small[0], small[1]);
            float csum = 0.f;
#pragma unroll
            for (int j = 0; j < KP; ++j) { small[2 + j] += coef[j]; csum += coef[j]; }
            if (a.dz) {                               // dz = (sum_j c_j) z - sum_j c_j mu_j, in centred form
                float dzr[D];
#pragma unroll
                for (int c = 0; c < D; ++c) dzr[c] = 0.f;
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    if (j < K) {
#pragma unroll
                        for (int c = 0; c < D; ++c) dzr[c] = fmaf(coef[j], zr[c] - mu_s[j * D + c], dzr[c]);
                    }
                }
                store_row<D>(out_tile, threadIdx.x, dzr);
            }
            // overwrite own row with the centred point for phase 2
#pragma unroll
            for (int c = 0; c < D; ++c) zr[c] -= c0_s[c];
            store_row<D>(ztile, threadIdx.x, zr);
        }
#pragma unroll
        for (int j = 0; j < KP; j += 4)
            *reinterpret_cast<float4*>(w_tile + threadIdx.x * KP + j) =
                make_float4(coef[j], coef[j + 1], coef[j + 2], coef[j + 3]);
        __syncthreads();
        if (a.dz) copy_tile_out<D, kDecThreads>(out_tile, a.dz + tile * (int64_t)kDecTile * D, np);
        if (p2_active) {
            // rows with index >= np carry zero coefficients; their z rows are stale but finite?  Not
            // guaranteed (uninitialised smem), so bound the loop by np.
            for (int r = warp * G + grp; r < np; r += NW * G) {
                const float4 w = *reinterpret_cast<const float4*>(w_tile + r * KP + 4 * jb);
                const float4 x = *reinterpret_cast<const float4*>(ztile + r * L::LD + 4 * cb);
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
        __syncthreads();                 // phase 2 done: stage and w_tile / out_tile reusable
        ring.issue(stage, tile + (int64_t)S * gridDim.x);
    }
    // ---- CTA reduction ----
    small[0] *= a.scale;
    cta_reduce<KP + 2, kDecThreads>(small, scratch, cta_stats + NS, 0);     // temp: [NS .. NS+KP+2) overlaps? sized below
    // cta_stats layout now: [NS + 0] loss, [NS + 1] sum s, [NS + 2 + j] Wsum_j.  Move loss/s to the front.
    // Per-(warp, group) 4x4 partials -> shared (reuse the ring buffer, free now), fixed-order sum.
    double* part = reinterpret_cast<double*>(ring_buf);                      // [NW*G][KP*D]
    if (p2_active) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                part[(size_t)(warp * G + grp) * (KP * D) + (4 * jb + r) * D + 4 * cb + c] = (double)blk[4 * r + c];
    }
    __syncthreads();
    const double loss = cta_stats[NS], ssum = cta_stats[NS + 1];
    __syncthreads();
    for (int o = threadIdx.x; o < KP * D; o += kDecThreads) {
        double accd = 0.0;
#pragma unroll
        for (int g = 0; g < NW * G; ++g) accd += part[(size_t)g * (KP * D) + o];
        const int j = o / D, c = o - j * D;
        const double wsum = cta_stats[NS + 2 + j];
        cta_stats[2 + o] = -(accd - wsum * ((double)mu_s[o] - (double)c0_s[c]));
    }
    if (threadIdx.x == 0) { cta_stats[0] = loss; cta_stats[1] = ssum; }
    __syncthreads();
    grid_publish(cta_stats, K * D + 2, a.partials, a.counter, a.stats);
}

// ---------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------
template <typename Kern>
static int launch_persistent(Kern kern, const DecArgs& args, size_t smem, int64_t num_tiles, cudaStream_t stream) {
    int64_t grid = persistent_grid(reinterpret_cast<const void*>(kern), kDecThreads, smem, kMaxCtasPerSm);
    if (grid < 0) return (int)grid;
    if (grid > kMaxDecGrid) grid = kMaxDecGrid;
    if (grid > num_tiles) grid = num_tiles;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kDecThreads, smem, stream>>>(args);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

template <int D, int KP>
static size_t assign_smem() {
    constexpr int S = dec_stages<D>();
    return sizeof(float) * (S * kDecTile * RowLayout<D>::LD + ((KP * D + 3) & ~3)) +
           sizeof(double) * ((kDecThreads / 32) * (KP + 1) + (KP + 1)) + sizeof(uint64_t) * S;
}
template <int D, int KP>
static size_t grad_reg_smem() {
    constexpr int S = dec_stages<D>();
    constexpr int NV = KP * D + 2;
    return sizeof(float) * ((S + 1) * kDecTile * RowLayout<D>::LD + ((KP * D + 3) & ~3) + ((KP + 3) & ~3)) +
           sizeof(double) * ((kDecThreads / 32) * NV + NV) + sizeof(uint64_t) * S;
}
template <int D, int KP>
static size_t grad_tiled_smem() {
    constexpr int S = 2;
    constexpr int NB = (KP / 4) * (D / 4);
    constexpr int G = (32 / NB) > 0 ? (32 / NB) : 1;
    constexpr int NW = kDecThreads / 32;
    size_t bytes = sizeof(float) * ((S + 1) * kDecTile * RowLayout<D>::LD + kDecTile * KP + KP * D + D + KP) +
                   sizeof(double) * (NW * (KP + 2) + (KP * D + 2 + KP + 2)) + sizeof(uint64_t) * S;
    // the ring buffer is reused for the [NW*G][KP*D] float64 partials at the end
    const size_t part = sizeof(double) * NW * G * KP * D;
    const size_t ring = sizeof(float) * S * kDecTile * RowLayout<D>::LD;
    if (part > ring) bytes += part - ring;
    return bytes;
}

template <int D, int KP>
struct DecOps {
    static int assign(const DecArgs& a, cudaStream_t st) {
        const int64_t tiles = (a.n + kDecTile - 1) / kDecTile;
        if (a.alpha == 1.0f) return launch_persistent(dec_assign_kernel<D, KP, true>, a, assign_smem<D, KP>(), tiles, st);
        return launch_persistent(dec_assign_kernel<D, KP, false>, a, assign_smem<D, KP>(), tiles, st);
    }
    template <int MODE>
    static int grad(const DecArgs& a, cudaStream_t st) {
        const int64_t tiles = (a.n + kDecTile - 1) / kDecTile;
        constexpr bool kTiled = (KP * D > 160);
        if constexpr (kTiled) {
            if constexpr (D % 4 == 0 && KP % 4 == 0) {
                if (a.alpha == 1.0f)
                    return launch_persistent(dec_grad_tiled_kernel<D, KP, true, MODE>, a, grad_tiled_smem<D, KP>(), tiles, st);
                return launch_persistent(dec_grad_tiled_kernel<D, KP, false, MODE>, a, grad_tiled_smem<D, KP>(), tiles, st);
            } else {
                return SCC_ERR_UNSUPPORTED;
            }
        } else {
            if (a.alpha == 1.0f)
                return launch_persistent(dec_grad_reg_kernel<D, KP, true, MODE>, a, grad_reg_smem<D, KP>(), tiles, st);
            return launch_persistent(dec_grad_reg_kernel<D, KP, false, MODE>, a, grad_reg_smem<D, KP>(), tiles, st);
        }
    }
};

static int pick_kp(int K) { return K <= 4 ? 4 : (K <= 8 ? 8 : 16); }

#define SCC_DEC_DISPATCH(D_, KP_, CALL)                                   \
    if (d == D_ && kp == KP_) return DecOps<D_, KP_>::CALL;

#define SCC_DEC_DISPATCH_ALL(CALL)                                        \
    SCC_FOR_EACH_DIM(SCC_DEC_DISPATCH_D, CALL)

#define SCC_DEC_DISPATCH_D(D_, CALL)                                      \
    SCC_DEC_DISPATCH(D_, 4, CALL) SCC_DEC_DISPATCH(D_, 8, CALL) SCC_DEC_DISPATCH(D_, 16, CALL)

bool dec_supported(int d, int K) {
    if (K < 1 || K > SCC_MAX_K) return false;
#define SCC_SUP(D_, X) if (d == D_) return true;
    SCC_FOR_EACH_DIM(SCC_SUP, 0)
#undef SCC_SUP
    return false;
}

static int check_common(const float* z, int64_t n, int d, const float* mu, int K, float alpha, double* stats,
                        void* ws, size_t ws_bytes) {
    if ((!z && n > 0) || !mu || !stats || n < 0 || !(alpha > 0.f)) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if (!dec_supported(d, K)) return SCC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(z) & 15u) != 0) return SCC_ERR_MISALIGNED;
    if (!ws || ws_bytes < workspace_bytes(d, K)) return SCC_ERR_WORKSPACE;
    return SCC_OK;
}

static void fill_reduction(DecArgs& a, void* ws) {
    a.counter = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader);
}

int dec_assign(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
               float* q, int32_t* labels, const int32_t* labels_prev, double* stats,
               void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = check_common(z, n, d, mu, K, alpha, stats, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (round_decimals != 0 && round_decimals != 5) return SCC_ERR_INVALID;
    if (q && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(q) & 15u)) return SCC_ERR_MISALIGNED;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K + 1), st)); return SCC_OK; }
    DecArgs a{};
    a.z = z; a.n = n; a.mu = mu; a.K = K; a.alpha = alpha; a.round5 = round_decimals == 5;
    a.q = q; a.labels = labels; a.labels_prev = labels_prev; a.stats = stats;
    fill_reduction(a, ws);
    const int kp = pick_kp(K);
    SCC_DEC_DISPATCH_ALL(assign(a, st))
    return SCC_ERR_UNSUPPORTED;
}

int dec_target(const float* q, int64_t n, int K, const double* f, int round_decimals, float* p, cudaStream_t st) {
    if (!q || !f || !p || n < 0 || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if (round_decimals != 0 && round_decimals != 5) return SCC_ERR_INVALID;
    if (n == 0) return SCC_OK;
    int dev = 0, sms = 0;
    SCC_CUDA(cudaGetDevice(&dev));
    SCC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool vec = (K % 4 == 0) && ((K / 4) == 1 || (K / 4) == 2 || (K / 4) == 4) &&
                     !((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(p)) & 15u);
    const int64_t work = vec ? n * (K / 4) : n;
    int64_t grid = (work + 255) / 256;
    const int64_t cap = (int64_t)sms * 8;
    if (grid > cap) grid = cap;
    const int r5 = round_decimals == 5;
    if (vec && K == 4) dec_target_kernel<1><<<(unsigned)grid, 256, 0, st>>>(q, n, K, f, r5, p);
    else if (vec && K == 8) dec_target_kernel<2><<<(unsigned)grid, 256, 0, st>>>(q, n, K, f, r5, p);
    else if (vec && K == 16) dec_target_kernel<4><<<(unsigned)grid, 256, 0, st>>>(q, n, K, f, r5, p);
    else dec_target_kernel<0><<<(unsigned)grid, 256, 0, st>>>(q, n, K, f, r5, p);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}


// ---------------------------------------------------------------------------
// colsum: f_j = sum_i q_ij for a caller-supplied q (models.py:1320)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecThreads)
colsum_kernel(const float* __restrict__ q, int64_t n, int K, double* stats, double* partials, unsigned int* counter) {
    constexpr int KP = SCC_MAX_K;
    __shared__ double scratch[(kDecThreads / 32) * KP];
    __shared__ double cta_stats[KP];
    float acc[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) acc[j] = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v[KP];
        load_krow<KP>(q + i * K, K, v);
#pragma unroll
        for (int j = 0; j < KP; ++j) acc[j] += v[j];
    }
    cta_reduce<KP, kDecThreads>(acc, scratch, cta_stats, 0);
    grid_publish(cta_stats, K, partials, counter, stats);
}

int colsum(const float* q, int64_t n, int K, double* f, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!q || !f || n < 0 || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if ((K % 4 == 0) && (reinterpret_cast<uintptr_t>(q) & 15u)) return SCC_ERR_MISALIGNED;
    if (!ws || ws_bytes < workspace_bytes(4, K)) return SCC_ERR_WORKSPACE;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(f, 0, sizeof(double) * K, st)); return SCC_OK; }
    int dev = 0, sms = 0;
    SCC_CUDA(cudaGetDevice(&dev));
    SCC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t grid = (n + kDecThreads - 1) / kDecThreads;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    if (grid > kMaxDecGrid) grid = kMaxDecGrid;
    colsum_kernel<<<(unsigned)grid, kDecThreads, 0, st>>>(
        q, n, K, f, reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader),
        reinterpret_cast<unsigned int*>(ws));
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

int dec_kl_grad(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* p,
                const double* f_cols, int round_decimals, float scale, float* dz, double* stats,
                void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = check_common(z, n, d, mu, K, alpha, stats, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (!p && !f_cols) return SCC_ERR_INVALID;
    if (round_decimals != 0 && round_decimals != 5) return SCC_ERR_INVALID;
    if (p && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(p) & 15u)) return SCC_ERR_MISALIGNED;
    if (dz && (reinterpret_cast<uintptr_t>(dz) & 15u)) return SCC_ERR_MISALIGNED;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K * d + 2), st)); return SCC_OK; }
    DecArgs a{};
    a.z = z; a.n = n; a.mu = mu; a.K = K; a.alpha = alpha; a.round5 = round_decimals == 5;
    a.p = p; a.f_cols = f_cols; a.scale = scale; a.dz = dz; a.stats = stats;
    fill_reduction(a, ws);
    const int kp = pick_kp(K);
    SCC_DEC_DISPATCH_ALL(template grad<MODE_KL>(a, st))
    return SCC_ERR_UNSUPPORTED;
}

int dec_backward(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* grad_q,
                 float* dz, double* stats, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = check_common(z, n, d, mu, K, alpha, stats, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (!grad_q) return SCC_ERR_INVALID;
    if ((K % 4 == 0) && (reinterpret_cast<uintptr_t>(grad_q) & 15u)) return SCC_ERR_MISALIGNED;
    if (dz && (reinterpret_cast<uintptr_t>(dz) & 15u)) return SCC_ERR_MISALIGNED;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K * d + 2), st)); return SCC_OK; }
    DecArgs a{};
    a.z = z; a.n = n; a.mu = mu; a.K = K; a.alpha = alpha;
    a.grad_q = grad_q; a.scale = 1.f; a.dz = dz; a.stats = stats;
    fill_reduction(a, ws);
    const int kp = pick_kp(K);
    SCC_DEC_DISPATCH_ALL(template grad<MODE_GENERIC>(a, st))
    return SCC_ERR_UNSUPPORTED;
}

}  // namespace scc
