// gmm_api.cu — GMM entry points: finalize / pack kernels (dimension independent, float64) and the
// per-dimension dispatch of the fused EM pass.
#include "gmm_kernels.cuh"

namespace scc {

#define SCC_FOR_EACH_GMM_DIM(M) M(4) M(8) M(9) M(10) M(12) M(16) M(20) M(24) M(32)
#define SCC_DECL(D_) int gmm_em_dim##D_(const GmmArgs& a, cudaStream_t st);
SCC_FOR_EACH_GMM_DIM(SCC_DECL)
#undef SCC_DECL

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
    unsigned long long* tl;    // profiling builds: timeline row of the tail kernel (slots 3..6), or NULL
};

// One component, one warp, the matrix in REGISTERS (d <= 12): lane i holds row i of the covariance / of L in
// a[0..D), lane c ends with column c of Y = L^-1 in y[c+1..D) and 1 / L[c][c] in ycc.  Right-looking Cholesky —
// after column j the trailing rows are updated with shuffled L[c][j] — and forward substitution with shuffled
// L[r][m]: the same products in the same order as the shared-memory version (bit-identical L and Y), without
// its dependent shared-memory round trips.  Returns false when a pivot is not positive (warp-uniform).
template <int D>
__device__ __forceinline__ bool chol_inverse_regs(double (&a)[D], double (&y)[D], double& ycc, int lane) {
    constexpr unsigned int kFull = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const double piv = __shfl_sync(kFull, a[j], j);
        if (!(piv > 0.0)) return false;
        const double rinv = rsqrt(piv);
        const double l = (lane == j) ? piv * rinv : a[j] * rinv;       // L[lane][j] (lanes < j: never read)
        a[j] = l;
#pragma unroll
        for (int c = j + 1; c < D; ++c) {
            const double lc = __shfl_sync(kFull, l, c);                // L[c][j]
            a[c] = fma(-l, lc, a[c]);
        }
    }
    double diag = 1.0;
#pragma unroll
    for (int m = 0; m < D; ++m) {
        if (m == lane) diag = a[m];
        y[m] = 0.0;
    }
    ycc = 1.0 / diag;
#pragma unroll
    for (int r = 1; r < D; ++r) {
        const double yrr = __shfl_sync(kFull, ycc, r);
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < r; ++m) {
            const double lrm = __shfl_sync(kFull, a[m], r);            // L[r][m]
            if (m == lane) acc = lrm * ycc;
            else if (m > lane) acc = fma(lrm, y[m], acc);
        }
        if (lane < r) y[r] = -acc * yrr;
    }
    return true;
}

// All threads of one CTA of at least 32 * K threads.
template <bool FROM_STATS, int DR>          // DR: compile-time d of the register path, 0 = any d through shared memory
__device__ __forceinline__ void gmm_finalize_body(const FinArgs& a, unsigned char* smem_raw) {
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

    if constexpr (DR > 0) {
        if (k < K) {
            constexpr int D = DR;
            const int i = lane;
            double row[D], y[D], ycc, mean_new = 0.0;
            if constexpr (FROM_STATS) {
                const double* N = a.stats + 1;
                const double* S1 = a.stats + 1 + K;
                const double* s2k = a.stats + 1 + K + (size_t)K * D + (size_t)k * TRI;
                // every load first (independent: one L2 round trip), one reciprocal instead of d + 1 divisions
                const double nk = N[k] + a.nk_add;
                const double s1 = (i < D) ? S1[k * D + i] : 0.0;
                const double mean_old = (i < D) ? a.means[k * D + i] : 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const int lo = i < c ? i : c, hi = i < c ? c : i;
                    row[c] = (i < D) ? s2k[tri(hi) + lo] : 0.0;
                }
                if (lane == 0) nk_s[k] = nk;
                const double rnk = 1.0 / nk;
                const double di = s1 * rnk;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const double dc = __shfl_sync(0xffffffffu, di, c);
                    double v = row[c] * rnk - di * dc;
                    if (c == i) v += a.reg_covar;
                    row[c] = v;
                    if (i < D) a.covariances[((size_t)k * D + i) * D + c] = v;
                }
                if (i < D) {
                    mean_new = mean_old + di;
                    a.means[k * D + i] = mean_new;
                }
            } else {
                if (lane == 0) nk_s[k] = a.weights_in[k];
#pragma unroll
                for (int c = 0; c < D; ++c) row[c] = (i < D) ? a.cov_in[((size_t)k * D + i) * D + c] : 0.0;
                if (i < D) mean_new = a.means[k * D + i];
            }
            if (k == 0) SCC_TL(a.tl, 3);
            const bool ok = chol_inverse_regs<D>(row, y, ycc, lane);
            if (k == 0) SCC_TL(a.tl, 5);
            if (!ok) {
                if (lane == 0) atomicMax(&bad_s, k + 1);
            } else {
                double ld = (i < D) ? log(ycc) : 0.0;                  // log det U = -sum log L_ii
                ld = warp_sum(ld);
                if (lane == 0) logdet_s[k] = ld;
                if (i < D) {
#pragma unroll
                    for (int b = 0; b < D; ++b) {
                        const double u = (i < b) ? y[b] : ((i == b) ? ycc : 0.0);       // U[i][b] = Y[b][i]
                        if (a.prec_chol) a.prec_chol[((size_t)k * D + i) * D + b] = u;
                        if (i <= b) a.params[(size_t)K * D + (size_t)k * TRI + tri(b) + i] = (float)u;
                    }
                    a.params[k * D + i] = (float)mean_new;
                }
            }
        }
    } else if (k < K) {
        double* A = mats + (size_t)k * d * LDA;
        const int i = lane;                         // row owned by this lane
        double mean_new = 0.0;                      // this lane's coordinate of the updated mean
        if constexpr (FROM_STATS) {
            const double* N = a.stats + 1;
            const double* S1 = a.stats + 1 + K;
            const double* S2 = a.stats + 1 + K + (size_t)K * d;
            const double nk = N[k] + a.nk_add;
            if (lane == 0) nk_s[k] = nk;
            // loads first (independent, nothing stored to global memory in between: they overlap instead of paying one
            // L2 round trip per matrix entry), then the arithmetic; the mean shift of the other rows comes by shuffle
            const double di = (i < d) ? S1[k * d + i] / nk : 0.0;
            if (i < d) {
                const double* __restrict__ s2k = S2 + (size_t)k * TRI;
                for (int c = 0; c < d; ++c) {
                    const int lo = i < c ? i : c, hi = i < c ? c : i;
                    A[i * LDA + c] = s2k[tri(hi) + lo];
                }
            }
            for (int c = 0; c < d; ++c) {
                const double dc = __shfl_sync(0xffffffffu, di, c);
                if (i < d) {
                    double v = A[i * LDA + c] / nk - di * dc;
                    if (c == i) v += a.reg_covar;
                    A[i * LDA + c] = v;
                    a.covariances[((size_t)k * d + i) * d + c] = v;
                }
            }
            if (i < d) {
                mean_new = a.means[k * d + i] + di;
                a.means[k * d + i] = mean_new;
            }
        } else {
            if (i < d) mean_new = a.means[k * d + i];
            if (lane == 0) nk_s[k] = a.weights_in[k];
            if (i < d)
                for (int c = 0; c < d; ++c) A[i * LDA + c] = a.cov_in[((size_t)k * d + i) * d + c];
        }
        __syncwarp();
        if (k == 0) SCC_TL(a.tl, 3);
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
            const double rinv = rsqrt(piv);                  // one dependent special-function chain per column, not two
            if (i == j) A[i * LDA + j] = piv * rinv;
            else if (i > j && i < d) A[i * LDA + j] = s * rinv;
            __syncwarp();
        }
        if (k == 0) SCC_TL(a.tl, 4);
        if (!ok) {
            if (lane == 0) atomicMax(&bad_s, k + 1);
        } else {
            // Y = L^-1 (lower triangular), lane c owns column c (forward substitution).  Y[r][c], r > c,
            // is parked in the strict UPPER triangle at A[c][r]; Y[c][c] = 1 / L[c][c].  The upper
            // triangle of A with the diagonal inverted is then exactly U = L^-T.
            const int c = lane;
            const double ycc = (c < d) ? 1.0 / A[c * LDA + c] : 1.0;   // the d reciprocals side by side, then shuffled
            for (int r = 1; r < d; ++r) {
                const double yrr = __shfl_sync(0xffffffffu, ycc, r);
                if (c < r) {
                    double acc = A[r * LDA + c] * ycc;
                    for (int m = c + 1; m < r; ++m) acc += A[r * LDA + m] * A[c * LDA + m];
                    A[c * LDA + r] = -acc * yrr;
                }
            }
            __syncwarp();
            if (k == 0) SCC_TL(a.tl, 5);
            double ld = (i < d) ? log(ycc) : 0.0;                      // log det U = -sum log L_ii
            ld = warp_sum(ld);
            if (lane == 0) logdet_s[k] = ld;
            if (i < d) {
                for (int b = 0; b < d; ++b) {
                    const double u = (i < b) ? A[i * LDA + b] : ((i == b) ? ycc : 0.0);   // U[i][b]
                    if (a.prec_chol) a.prec_chol[((size_t)k * d + i) * d + b] = u;
                    if (i <= b) a.params[(size_t)K * d + (size_t)k * TRI + tri(b) + i] = (float)u;
                }
                a.params[k * d + i] = (float)mean_new;
            }
        }
    }
    __syncthreads();
    SCC_TL(a.tl, 6);
    if ((int)threadIdx.x < K && bad_s == 0) {          // one lane per component: K float64 logarithms side by side
        double tot = 0.0;
        for (int j = 0; j < K; ++j) tot += nk_s[j];   // same order in every lane
        const int j = threadIdx.x;
        const double w = nk_s[j] / tot;
        if (a.weights) a.weights[j] = w;
        a.params[(size_t)K * d + (size_t)K * TRI + j] =
            (float)(logdet_s[j] + log(w) - 0.5 * d * 1.8378770664093453 /* log(2 pi) */);
    }
    if (threadIdx.x == 0) {
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

// d <= 12 of the built dimensions take the register path
template <bool FROM_STATS>
__device__ __forceinline__ void gmm_finalize_dispatch(const FinArgs& a, unsigned char* smem_raw) {
    switch (a.d) {
        case 4: gmm_finalize_body<FROM_STATS, 4>(a, smem_raw); break;
        case 8: gmm_finalize_body<FROM_STATS, 8>(a, smem_raw); break;
        case 9: gmm_finalize_body<FROM_STATS, 9>(a, smem_raw); break;
        case 10: gmm_finalize_body<FROM_STATS, 10>(a, smem_raw); break;
        case 12: gmm_finalize_body<FROM_STATS, 12>(a, smem_raw); break;
        default: gmm_finalize_body<FROM_STATS, 0>(a, smem_raw); break;
    }
}

template <bool FROM_STATS>
__global__ void __launch_bounds__(512, 1)
gmm_finalize_kernel(const FinArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    gmm_finalize_dispatch<FROM_STATS>(a, smem_raw);
}

// ---------------------------------------------------------------------------
// Tail of one EM iteration in ONE launch: fixed-order reduction of the per-CTA partial statistics, their
// cross-GPU all-reduce and the M-step finalisation.  Every CTA owns a slice of the statistics vector: it sums
// the slice over the statistics kernel's partial slots, ships it to every rank's exchange window and polls the
// same slice of every rank (flag-in-data: slices are independent), writes the world's sums to `stats`; the CTA
// that finishes last then runs the finalisation (scaling, Cholesky, U = L^-T, lower bound, stop rule).
// Replaces reduce_partials -> exchange -> finalize (three launches) behind the statistics kernel.
// ---------------------------------------------------------------------------
constexpr int kTailSlice = 64;             // statistics per CTA: 512 threads = 64 statistics x 8 groups of partial slots
constexpr int kTailGroups = 8;

__global__ void __launch_bounds__(512, 1)
gmm_tail_kernel(const double* __restrict__ partials, int G, int NS, double* __restrict__ stats, PeerCtx ex,
                unsigned int* __restrict__ ticket, const FinArgs fin, unsigned long long* tl) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double slice[kTailSlice];
    __shared__ double part[kTailGroups][kTailSlice];
    __shared__ int s_last;
    pdl_wait();                                  // launched under the statistics kernel's drain: no global access before
    if (fin.ctrl[5] != 0.0) return;              // frozen fit (identical on every rank): nothing to exchange
    SCC_TL(tl, 0);
    const int lo = blockIdx.x * kTailSlice, hi = min(NS, lo + kTailSlice);
    {   // thread (g, c): statistic lo + c over the slots g, g + 8, ... — the chain is L2-latency bound, so the slots
        // of one statistic are spread over 8 threads (up to 8 independent loads in flight each); fixed order
        const int c = threadIdx.x % kTailSlice, g = threadIdx.x / kTailSlice;
        const int s = lo + c;
        double acc = 0.0;
        if (s < hi) {
            for (int b = g; b < G; b += kTailGroups * 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int bb = b + u * kTailGroups;
                    v[u] = (bb < G) ? partials[(size_t)bb * NS + s] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += v[u];
            }
        }
        part[g][c] = acc;
    }
    __syncthreads();
    if ((int)threadIdx.x < kTailSlice) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < kTailGroups; ++g) t += part[g][threadIdx.x];
        slice[threadIdx.x] = t;
    }
    __syncthreads();
    const int s = lo + threadIdx.x;
    if (ex.windows) {
        PeerHeader* me = reinterpret_cast<PeerHeader*>(ex.windows[ex.rank]);
        const unsigned int seq = ld_relaxed_gpu_u32(&me->seq) + 1u;
        peer_push_slice(ex, slice, lo, hi, seq);
        peer_pull_slice(ex, stats + lo, lo, hi, seq);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int t;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(&me->ticket) : "memory");
            if (t == gridDim.x - 1) {
                me->ticket = 0u;
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(&me->seq), "r"(seq) : "memory");
            }
        }
    } else if (s < hi) {
        stats[s] = slice[threadIdx.x];
    }
    static_assert(kTailSlice * kTailGroups == 512, "tail kernel: 512 threads");
    // the CTA that finishes last sees every slice of `stats` (writes above -> CTA barrier -> acq_rel ticket)
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
        s_last = (t == gridDim.x - 1) ? 1 : 0;
        if (s_last) *ticket = 0u;
    }
    __syncthreads();
    SCC_TL(tl, 1);
    if (!s_last) return;
    gmm_finalize_dispatch<true>(fin, smem_raw);
    SCC_TL(tl, 2);
}


bool gmm_supported(int d, int K) {
    if (K < 1 || K > SCC_MAX_K) return false;
#define SCC_SUP(D_) if (d == D_) return true;
    SCC_FOR_EACH_GMM_DIM(SCC_SUP)
#undef SCC_SUP
    return false;
}

int gmm_em_step(const float* z, int64_t n, int d, int K, const float* params, double* stats,
                int32_t* labels, float* resp, const double* ctrl, int mode,
                void* ws, size_t ws_bytes, cudaStream_t st) {
    if ((!z && n > 0) || !params || !stats || n < 0) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if (mode < 0 || (mode & 3) > 2 || (mode & ~(3 | SCC_GMM_NOSKIP))) return SCC_ERR_INVALID;
    if (!gmm_supported(d, K)) return SCC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(z) & 15u) != 0) return SCC_ERR_MISALIGNED;
    if (!ws || ws_bytes < workspace_bytes(d, K)) return SCC_ERR_WORKSPACE;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * SCC_GMM_STAT_DOUBLES(K, d), st)); return SCC_OK; }
    GmmArgs a{};
    a.z = z; a.n = n; a.K = K; a.params = params; a.labels = labels; a.resp = resp; a.ctrl = ctrl;
    a.accumulate = mode; a.stats = stats; a.timeline = g_timeline;
    a.counter = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader);
#define SCC_CASE(D_) if (d == D_) return gmm_em_dim##D_(a, st);
    SCC_FOR_EACH_GMM_DIM(SCC_CASE)
#undef SCC_CASE
    return SCC_ERR_UNSUPPORTED;
}

static size_t finalize_smem(int d, int K) { return sizeof(double) * (size_t)K * d * (d + 1); }

int gmm_em_iteration(const float* z, int64_t n, int d, int K, float* params, double* stats, int mode, double n_total,
                     double reg_covar, double nk_eps, double tol, double* means, double* weights, double* covariances,
                     double* prec_chol, double* ctrl, void* ws, size_t ws_bytes, const ExchangeDesc* ex, cudaStream_t st) {
    if ((!z && n > 0) || !params || !stats || !means || !weights || !covariances || !ctrl || n < 0) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K || !(n_total > 0)) return SCC_ERR_INVALID;
    if ((mode & 3) == 0 || (mode & 3) > 2 || (mode & ~(3 | SCC_GMM_NOSKIP))) return SCC_ERR_INVALID;
    if (!gmm_supported(d, K)) return SCC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(z) & 15u) != 0) return SCC_ERR_MISALIGNED;
    if (!ws || ws_bytes < workspace_bytes(d, K)) return SCC_ERR_WORKSPACE;
    const int NS = SCC_GMM_STAT_DOUBLES(K, d);
    if (ex && ex->windows && NS > ex->max_len) return SCC_ERR_INVALID;
    int grid = 0;
    GmmArgs a{};
    a.z = z; a.n = n; a.K = K; a.params = params; a.ctrl = ctrl; a.accumulate = mode; a.stats = stats;
    a.counter = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader);
    a.skip_reduce = 1; a.grid_out = &grid; a.timeline = g_timeline;
    if (n > 0) {
        int rc = SCC_ERR_UNSUPPORTED;
#define SCC_CASE(D_) if (d == D_) rc = gmm_em_dim##D_(a, st);
        SCC_FOR_EACH_GMM_DIM(SCC_CASE)
#undef SCC_CASE
        if (rc != SCC_OK) return rc;
    }
    FinArgs f{};
    f.stats = stats; f.n_total = n_total; f.reg_covar = reg_covar; f.nk_add = nk_eps; f.tol = tol;
    f.d = d; f.K = K; f.means = means; f.weights = weights; f.covariances = covariances;
    f.prec_chol = prec_chol; f.params = params; f.ctrl = ctrl;
    f.tl = g_timeline ? g_timeline + 8 * 1024 : nullptr;
    PeerCtx px{nullptr, 0, 1, 0};
    if (ex && ex->windows) px = PeerCtx{reinterpret_cast<unsigned char* const*>(ex->windows), ex->rank, ex->world, ex->max_len};
    const size_t smem = finalize_smem(d, K);
    SCC_CUDA(cudaFuncSetAttribute(gmm_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // programmatic dependent launch: the tail's CTAs are scheduled as the statistics kernel's CTAs retire and wait in
    // griddepcontrol.wait for the whole grid (the launch gap, ~2.5 us per iteration, disappears under the drain)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((NS + kTailSlice - 1) / kTailSlice));
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    unsigned long long* tl = g_timeline ? g_timeline + 8 * 1024 : nullptr;
    SCC_CUDA(cudaLaunchKernelEx(&cfg, gmm_tail_kernel, (const double*)a.partials, grid, NS, stats, px, a.counter + 4, f, tl));
    return SCC_OK;
}

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
