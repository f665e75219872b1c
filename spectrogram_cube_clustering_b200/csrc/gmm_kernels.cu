// gmm_kernels.cu — full-covariance Gaussian-mixture EM for sm_100a (FP32 CUDA cores).
//
//   gmm_em_full_kernel   fused E-step + M-step statistics, one read of z      (d <= 12)
//       phase 1: one thread per point — Cholesky log-likelihoods for all K components
//                (sklearn _gaussian_mixture.py:490-553), log-sum-exp responsibilities
//                (_base.py:552-582), labels; r_ik parked in shared memory
//       phase 2: one warp per component — the lanes sweep the tile's points and keep the
//                1 + d + d(d+1)/2 moments of "their" component in registers, centred on
//                the current mean (sklearn _gaussian_mixture.py:282-320,168-197)
//   gmm_finalize_kernel  N_k, pi, mu, Sigma(+reg), Cholesky, U = L^-T, log det,
//                lower bound + convergence flag — one CTA, one warp per component, float64
//                (sklearn _gaussian_mixture.py:883-901,323-385,448-487; _base.py:270-278)
//
// This stage is FP32-FMA-bound, not HBM-bound: ~K(d^2+4d) FMA per point
// against 4d bytes.
#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

constexpr int kGmmFullMaxD = 12;

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
};

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

// ---------------------------------------------------------------------------
// finalize / pack: one CTA, warp k owns component k.  float64 throughout.
// ---------------------------------------------------------------------------
struct FinArgs {
    const double* stats;       // FROM_STATS
    const double* weights_in;  // !FROM_STATS
    const double* cov_in;      // !FROM_STATS
    double n_total, reg_covar, nk_add, tol;
    int d, K;
    double* means;             // in/out
    double* weights;           // out (may be NULL in pack mode)
    double* covariances;       // out (may be NULL in pack mode)
    double* prec_chol;         // out, may be NULL
    float* params;             // out
    double* ctrl;
};

template <bool FROM_STATS>
__global__ void __launch_bounds__(512, 1)
gmm_finalize_kernel(const FinArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = a.d, K = a.K, LDA = d + 1, TRI = tri(d);
    double* mats = reinterpret_cast<double*>(smem_raw);       // [K][d*LDA]  L below, Y^T above
    __shared__ double nk_s[SCC_MAX_K];
    __shared__ double logdet_s[SCC_MAX_K];
    __shared__ int bad_s;
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    double* ctrl = a.ctrl;
    if (FROM_STATS && ctrl[5] != 0.0) return;
    if (threadIdx.x == 0) bad_s = 0;
    __syncthreads();

    if (k < K) {
        double* A = mats + (size_t)k * d * LDA;
        const int i = lane;                         // row owned by this lane
        if constexpr (FROM_STATS) {
            const double* N = a.stats + 1;
            const double* S1 = a.stats + 1 + K;
            const double* S2 = a.stats + 1 + K + (size_t)K * d;
            const double nk = N[k] + a.nk_add;
            if (lane == 0) nk_s[k] = nk;
            if (i < d) {
                const double di = S1[k * d + i] / nk;
                for (int c = 0; c < d; ++c) {
                    const double dc = S1[k * d + c] / nk;
                    const int lo = i < c ? i : c, hi = i < c ? c : i;
                    double v = S2[(size_t)k * TRI + tri(hi) + lo] / nk - di * dc;
                    if (c == i) v += a.reg_covar;
                    A[i * LDA + c] = v;
                    a.covariances[((size_t)k * d + i) * d + c] = v;
                }
                a.means[k * d + i] += di;
            }
        } else {
            if (lane == 0) nk_s[k] = a.weights_in[k];
            if (i < d)
                for (int c = 0; c < d; ++c) A[i * LDA + c] = a.cov_in[((size_t)k * d + i) * d + c];
        }
        __syncwarp();
        // Cholesky, lower, in place (left-looking; lane i owns row i)
        bool ok = true;
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            if (i >= j && i < d) {
                s = A[i * LDA + j];
                for (int c = 0; c < j; ++c) s -= A[i * LDA + c] * A[j * LDA + c];
            }
            const double piv = __shfl_sync(0xffffffffu, s, j);
            if (!(piv > 0.0)) { ok = false; break; }
            const double root = sqrt(piv);
            if (i == j) A[i * LDA + j] = root;
            else if (i > j && i < d) A[i * LDA + j] = s / root;
            __syncwarp();
        }
        if (!ok) {
            if (lane == 0) atomicMax(&bad_s, k + 1);
        } else {
            // Y = L^-1 (lower triangular), lane c owns column c (forward substitution).  Y[r][c], r > c,
            // is parked in the strict UPPER triangle at A[c][r]; Y[c][c] = 1 / L[c][c].  The upper
            // triangle of A with the diagonal inverted is then exactly U = L^-T.
            const int c = lane;
            if (c < d) {
                const double ycc = 1.0 / A[c * LDA + c];
                for (int r = c + 1; r < d; ++r) {
                    double acc = A[r * LDA + c] * ycc;
                    for (int m = c + 1; m < r; ++m) acc += A[r * LDA + m] * A[c * LDA + m];
                    A[c * LDA + r] = -acc / A[r * LDA + r];
                }
            }
            __syncwarp();
            double ld = (i < d) ? -log(A[i * LDA + i]) : 0.0;          // log det U = -sum log L_ii
            ld = warp_sum(ld);
            if (lane == 0) logdet_s[k] = ld;
            if (i < d) {
                for (int b = 0; b < d; ++b) {
                    const double u = (i < b) ? A[i * LDA + b] : ((i == b) ? 1.0 / A[i * LDA + i] : 0.0);   // U[i][b]
                    if (a.prec_chol) a.prec_chol[((size_t)k * d + i) * d + b] = u;
                    if (i <= b) a.params[(size_t)K * d + (size_t)k * TRI + tri(b) + i] = (float)u;
                }
                a.params[k * d + i] = (float)a.means[k * d + i];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int j = 0; j < K; ++j) tot += nk_s[j];
        if (bad_s == 0) {
            for (int j = 0; j < K; ++j) {
                const double w = nk_s[j] / tot;
                if (a.weights) a.weights[j] = w;
                a.params[(size_t)K * d + (size_t)K * TRI + j] =
                    (float)(logdet_s[j] + log(w) - 0.5 * d * 1.8378770664093453 /* log(2 pi) */);
            }
        }
        if constexpr (FROM_STATS) {
            const double lower = a.stats[0] / a.n_total;
            const double prev = ctrl[0];
            ctrl[1] = prev; ctrl[0] = lower; ctrl[2] += 1.0;
            if (bad_s) { ctrl[4] = (double)bad_s; ctrl[5] = 1.0; }
            else if (fabs(lower - prev) < a.tol) { ctrl[3] = 1.0; ctrl[5] = 1.0; }
        } else {
            ctrl[0] = -INFINITY; ctrl[1] = -INFINITY; ctrl[2] = 0.0; ctrl[3] = 0.0;
            ctrl[4] = (double)bad_s; ctrl[5] = bad_s ? 1.0 : 0.0; ctrl[6] = 0.0; ctrl[7] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------
template <int D, int KP>
static size_t gmm_full_smem() {
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
    const int NS = SCC_GMM_STAT_DOUBLES(a.K, D);
    reduce_partials_kernel<<<(NS + 255) / 256, 256, 0, st>>>(a.partials, NS, (int)grid, a.stats, a.ctrl);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

bool gmm_supported(int d, int K) {
    if (K < 1 || K > SCC_MAX_K) return false;
    return d == 4 || d == 8 || d == 9 || d == 10 || d == 12;
}

int gmm_em_step(const float* z, int64_t n, int d, int K, const float* params, double* stats,
                int32_t* labels, float* resp, const double* ctrl, int mode,
                void* ws, size_t ws_bytes, cudaStream_t st) {
    if ((!z && n > 0) || !params || !stats || n < 0) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if (!gmm_supported(d, K)) return SCC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(z) & 15u) != 0) return SCC_ERR_MISALIGNED;
    if (!ws || ws_bytes < workspace_bytes(d, K)) return SCC_ERR_WORKSPACE;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * SCC_GMM_STAT_DOUBLES(K, d), st)); return SCC_OK; }
    GmmArgs a{};
    a.z = z; a.n = n; a.K = K; a.params = params; a.labels = labels; a.resp = resp; a.ctrl = ctrl;
    if (mode < 0 || mode > 2) return SCC_ERR_INVALID;
    a.accumulate = mode; a.stats = stats;
    a.counter = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader);
    const int kp = K <= 4 ? 4 : (K <= 8 ? 8 : 16);
#define SCC_GMM_CASE(D_, KP_) if (d == D_ && kp == KP_) return launch_gmm_full<D_, KP_>(a, st);
#define SCC_GMM_CASES(D_) SCC_GMM_CASE(D_, 4) SCC_GMM_CASE(D_, 8) SCC_GMM_CASE(D_, 16)
    SCC_GMM_CASES(4) SCC_GMM_CASES(8) SCC_GMM_CASES(9) SCC_GMM_CASES(10) SCC_GMM_CASES(12)
#undef SCC_GMM_CASES
#undef SCC_GMM_CASE
    return SCC_ERR_UNSUPPORTED;
}

static size_t finalize_smem(int d, int K) { return sizeof(double) * (size_t)K * d * (d + 1); }

int gmm_finalize(const double* stats, double n_total, int d, int K, double reg_covar, double nk_eps, double tol,
                 double* means, double* weights, double* covariances, double* prec_chol, float* params,
                 double* ctrl, cudaStream_t st) {
    if (!stats || !means || !weights || !covariances || !params || !ctrl) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K || !(n_total > 0)) return SCC_ERR_INVALID;
    FinArgs a{};
    a.stats = stats; a.n_total = n_total; a.reg_covar = reg_covar; a.nk_add = nk_eps; a.tol = tol;
    a.d = d; a.K = K; a.means = means; a.weights = weights; a.covariances = covariances;
    a.prec_chol = prec_chol; a.params = params; a.ctrl = ctrl;
    const size_t smem = finalize_smem(d, K);
    SCC_CUDA(cudaFuncSetAttribute(gmm_finalize_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_finalize_kernel<true><<<1, 512, smem, st>>>(a);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

int gmm_pack_params(const double* weights, const double* means, const double* covariances, int d, int K,
                    double* prec_chol, float* params, double* ctrl, cudaStream_t st) {
    if (!weights || !means || !covariances || !params || !ctrl) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    FinArgs a{};
    a.weights_in = weights; a.cov_in = covariances; a.d = d; a.K = K;
    a.means = const_cast<double*>(means); a.prec_chol = prec_chol; a.params = params; a.ctrl = ctrl;
    const size_t smem = finalize_smem(d, K);
    SCC_CUDA(cudaFuncSetAttribute(gmm_finalize_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmm_finalize_kernel<false><<<1, 512, smem, st>>>(a);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

}  // namespace scc
