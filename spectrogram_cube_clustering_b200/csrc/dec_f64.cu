// dec_f64.cu — float64 variant of the DEC clustering-layer path.
//
// The reference runs its DEC stage in float64 (`model.double()`, Cluster/models.py:965; numpy float64 in
// batch_eval / target_distribution, models.py:66-71, 1320-1322).  The float32 kernels of dec_kernels.cuh are the
// throughput path BASELINE.json asks for; these kernels are the PRECISION path: same operators, IEEE float64, the
// reference's own operation order, so that a caller who keeps the reference's dtype gets the reference's numbers
// (q, p to ~1e-15; the 5-decimal roundings land on the same side).  One thread per latent point, any d <= 32,
// K <= 16; centroids in shared memory; per-cluster sums reduced warp (shuffle) -> CTA (fixed warp order) -> grid
// (grid_publish: fixed slot order) — deterministic.  B200's FP64 rate is a fraction of its FP32 rate and rows are
// read straight from global memory: ~10x slower than the float32 path, still three orders of magnitude above the
// reference's CPU chain.
#include "scc_common.cuh"
#include "scc_launch.h"

namespace scc {

constexpr int kF64Threads = 256;

__device__ __forceinline__ double round5_f64(double x) { return rint(x * 100000.0) / 100000.0; }     // np.round(x, 5)

// q row of one point, reference operation order (networks.py:280-287).  Returns the arg max (first index).
__device__ __forceinline__ int soft_assign_f64(const double* __restrict__ zrow, const double* __restrict__ mu_s, int d, int K,
                                               double alpha, double expo, bool alpha1, double (&q)[SCC_MAX_K],
                                               double (&u)[SCC_MAX_K]) {
    double tsum = 0.0;
#pragma unroll
    for (int j = 0; j < SCC_MAX_K; ++j) {
        q[j] = 0.0; u[j] = 0.0;
        if (j < K) {
            double d2 = 0.0;
            for (int c = 0; c < d; ++c) {
                const double df = zrow[c] - mu_s[j * d + c];
                d2 += df * df;                                  // :281-282 (separate multiply and add)
            }
            const double x = 1.0 / (1.0 + d2 / alpha);          // :283-284
            u[j] = x;
            const double t = alpha1 ? x : pow(x, expo);         // :285
            q[j] = t;
            tsum += t;
        }
    }
    int label = 0;
    double best = -1.0;
#pragma unroll
    for (int j = 0; j < SCC_MAX_K; ++j) {
        if (j < K) {
            q[j] = q[j] / tsum;                                 // :286
            if (q[j] > best) { best = q[j]; label = j; }        // models.py:92 (first maximum)
        }
    }
    return label;
}

// CTA reduction of NV per-thread doubles (NV <= SCC_MAX_K + 2): shuffle tree per warp, fixed warp order.
__device__ __forceinline__ void cta_sum_f64(const double* vals, int nv, double* warp_s /*[8][nv]*/, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = 0; s < nv; ++s) {
        const double w = warp_sum(vals[s]);
        if (lane == 0) warp_s[warp * nv + s] = w;
    }
    __syncthreads();
    for (int s = threadIdx.x; s < nv; s += kF64Threads) {
        double acc = 0.0;
        for (int w = 0; w < kF64Threads / 32; ++w) acc += warp_s[w * nv + s];
        out[s] = acc;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// assign: q (optionally rounded), labels, f_j over the (rounded) q, label changes
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kF64Threads)
dec_assign_f64_kernel(const double* __restrict__ z, int64_t n, int d, const double* __restrict__ mu, int K, double alpha,
                      int round5, double* __restrict__ q_out, int32_t* __restrict__ labels,
                      const int32_t* __restrict__ labels_prev, double* stats, double* partials, unsigned int* counter) {
    __shared__ double mu_s[SCC_MAX_K * SCC_MAX_D];
    __shared__ double warp_s[(kF64Threads / 32) * (SCC_MAX_K + 1)];
    __shared__ double cta_stats[SCC_MAX_K + 1];
    __shared__ double scratch[2 * kF64Threads];
    for (int i = threadIdx.x; i < K * d; i += kF64Threads) mu_s[i] = mu[i];
    __syncthreads();
    const bool alpha1 = alpha == 1.0;
    const double expo = (alpha + 1.0) / 2.0;
    double facc[SCC_MAX_K + 1];
#pragma unroll
    for (int j = 0; j <= SCC_MAX_K; ++j) facc[j] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kF64Threads;
    for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += stride) {
        double zr[SCC_MAX_D];
        for (int c = 0; c < d; ++c) zr[c] = z[i * d + c];
        double q[SCC_MAX_K], u[SCC_MAX_K];
        const int label = soft_assign_f64(zr, mu_s, d, K, alpha, expo, alpha1, q, u);
#pragma unroll
        for (int j = 0; j < SCC_MAX_K; ++j) {
            if (j < K) {
                if (round5) q[j] = round5_f64(q[j]);                       // models.py:94
                facc[j] += q[j];
                if (q_out) q_out[i * K + j] = q[j];
            }
        }
        if (labels) labels[i] = label;
        if (labels_prev) facc[SCC_MAX_K] += (labels_prev[i] != label) ? 1.0 : 0.0;
    }
    double vals[SCC_MAX_K + 1];
#pragma unroll
    for (int j = 0; j < SCC_MAX_K; ++j) vals[j] = facc[j];
    vals[K] = facc[SCC_MAX_K];                                            // compact: [f[K], changed]
    cta_sum_f64(vals, K + 1, warp_s, cta_stats);
    grid_publish<kF64Threads>(cta_stats, K + 1, partials, counter, stats, scratch);
}

// ---------------------------------------------------------------------------
// column sums of a caller-supplied q (models.py:1320) and the target distribution (models.py:1320-1322)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kF64Threads)
colsum_f64_kernel(const double* __restrict__ q, int64_t n, int K, double* stats, double* partials, unsigned int* counter) {
    __shared__ double warp_s[(kF64Threads / 32) * SCC_MAX_K];
    __shared__ double cta_stats[SCC_MAX_K];
    __shared__ double scratch[2 * kF64Threads];
    double acc[SCC_MAX_K];
#pragma unroll
    for (int j = 0; j < SCC_MAX_K; ++j) acc[j] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * kF64Threads;
    for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += stride) {
#pragma unroll
        for (int j = 0; j < SCC_MAX_K; ++j)
            if (j < K) acc[j] += q[i * K + j];
    }
    cta_sum_f64(acc, K, warp_s, cta_stats);
    grid_publish<kF64Threads>(cta_stats, K, partials, counter, stats, scratch);
}

__global__ void __launch_bounds__(kF64Threads)
dec_target_f64_kernel(const double* __restrict__ q, int64_t n, int K, const double* __restrict__ f, int round5,
                      double* __restrict__ p) {
    const int64_t stride = (int64_t)gridDim.x * kF64Threads;
    for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n; i += stride) {
        double w[SCC_MAX_K];
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < SCC_MAX_K; ++j) {
            w[j] = 0.0;
            if (j < K) { const double x = q[i * K + j]; w[j] = x * x / f[j]; s += w[j]; }       // :1320
        }
#pragma unroll
        for (int j = 0; j < SCC_MAX_K; ++j) {
            if (j < K) {
                double v = w[j] / s;                                                             // :1321
                if (round5) v = round5_f64(v);                                                   // :1322
                p[i * K + j] = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// gradients.  mode 0: KL, target p streamed;  1: KL, p rebuilt from the column sums f (optionally written out);
//             2: generic upstream gradient dL/dq.
// stats out: [loss, sum_i s_i, dmu[K*d]].  Closed forms of SURVEY.md 8 a3 (oracle/dec.py kl_grads / backward_generic).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kF64Threads)
dec_grad_f64_kernel(const double* __restrict__ z, int64_t n, int d, const double* __restrict__ mu, int K, double alpha,
                    int mode, const double* __restrict__ p_in, const double* __restrict__ f, int round5,
                    const double* __restrict__ grad_q, double scale, double* __restrict__ p_out, double* __restrict__ dz,
                    double* stats, double* partials, unsigned int* counter) {
    extern __shared__ __align__(16) unsigned char f64_smem[];
    double* mu_s = reinterpret_cast<double*>(f64_smem);                   // [K*d]
    double* acc_s = mu_s + K * d;                                        // [8 warps][K*d + 2]
    double* cta_stats = acc_s + (kF64Threads / 32) * (K * d + 2);        // [K*d + 2]
    double* scratch = cta_stats + (K * d + 2);                           // [2 * threads]
    const int NS = K * d + 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < K * d; i += kF64Threads) mu_s[i] = mu[i];
    for (int i = threadIdx.x; i < (kF64Threads / 32) * NS; i += kF64Threads) acc_s[i] = 0.0;
    __syncthreads();
    const bool alpha1 = alpha == 1.0;
    const double expo = (alpha + 1.0) / 2.0;
    const double cs = (mode == 2 ? 1.0 : scale) * (alpha + 1.0) / alpha;
    double* mine = acc_s + warp * NS;
    const int64_t stride = (int64_t)gridDim.x * kF64Threads;
    const int64_t n_pad = (n + 31) & ~int64_t(31);                       // whole warps take part in the shuffles
    for (int64_t i = (int64_t)blockIdx.x * kF64Threads + threadIdx.x; i < n_pad; i += stride) {
        const bool active = i < n;
        double zr[SCC_MAX_D], q[SCC_MAX_K], u[SCC_MAX_K], w[SCC_MAX_K];
        double loss = 0.0, ssum = 0.0;
#pragma unroll
        for (int j = 0; j < SCC_MAX_K; ++j) w[j] = 0.0;
        if (active) {
            for (int c = 0; c < d; ++c) zr[c] = z[i * d + c];
            soft_assign_f64(zr, mu_s, d, K, alpha, expo, alpha1, q, u);
            if (mode == 2) {
                double dot = 0.0;
#pragma unroll
                for (int j = 0; j < SCC_MAX_K; ++j)
                    if (j < K) dot += grad_q[i * K + j] * q[j];
#pragma unroll
                for (int j = 0; j < SCC_MAX_K; ++j)
                    if (j < K) w[j] = -(q[j] * (grad_q[i * K + j] - dot)) * u[j] * cs;
            } else {
                double p[SCC_MAX_K];
                if (mode == 0) {
#pragma unroll
                    for (int j = 0; j < SCC_MAX_K; ++j) p[j] = (j < K) ? p_in[i * K + j] : 0.0;
                } else {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < SCC_MAX_K; ++j) {
                        p[j] = 0.0;
                        if (j < K) {
                            const double qr = round5 ? round5_f64(q[j]) : q[j];
                            p[j] = qr * qr / f[j];
                            s += p[j];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < SCC_MAX_K; ++j) {
                        if (j < K) {
                            p[j] = p[j] / s;
                            if (round5) p[j] = round5_f64(p[j]);
                            if (p_out) p_out[i * K + j] = p[j];
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < SCC_MAX_K; ++j) {
                    if (j < K) {
                        ssum += p[j];
                        if (p[j] > 0.0) loss += p[j] * (log(p[j]) - log(q[j]));       // torch KLDivLoss: xlogy
                        else if (!(p[j] == 0.0)) loss += p[j] * (log(p[j]) - log(q[j]));   // NaN / negative targets propagate
                    }
                }
#pragma unroll
                for (int j = 0; j < SCC_MAX_K; ++j)
                    if (j < K) w[j] = (p[j] - q[j] * ssum) * u[j] * cs;
            }
            if (dz) {
                for (int c = 0; c < d; ++c) {
                    double g = 0.0;
#pragma unroll
                    for (int j = 0; j < SCC_MAX_K; ++j)
                        if (j < K) g += w[j] * (zr[c] - mu_s[j * d + c]);
                    dz[i * d + c] = g;
                }
            }
        }
        // dmu_jc = -sum_i w_ij (z_ic - mu_jc): shuffle tree per (j, c), lane 0 adds into the warp's own accumulators
        {
            const double ls = warp_sum(active ? loss : 0.0), ss = warp_sum(active ? ssum : 0.0);
            if (lane == 0) { mine[0] += ls; mine[1] += ss; }
        }
#pragma unroll
        for (int j = 0; j < SCC_MAX_K; ++j) {
            if (j < K) {
                for (int c = 0; c < d; ++c) {
                    const double v = warp_sum(active ? -w[j] * (zr[c] - mu_s[j * d + c]) : 0.0);
                    if (lane == 0) mine[2 + j * d + c] += v;
                }
            }
        }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < NS; s += kF64Threads) {
        double a = 0.0;
        for (int w = 0; w < kF64Threads / 32; ++w) a += acc_s[w * NS + s];
        cta_stats[s] = (s == 0 && mode != 2) ? a * scale : a;
    }
    __syncthreads();
    grid_publish<kF64Threads>(cta_stats, NS, partials, counter, stats, scratch);
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static int f64_grid(int64_t n) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = (n + kF64Threads - 1) / kF64Threads;
    if (grid > (int64_t)sms * 4) grid = (int64_t)sms * 4;
    return grid < 1 ? 1 : (int)grid;
}

static int check_f64(const void* z, int64_t n, int d, const void* mu, int K, double alpha, void* ws, size_t ws_bytes) {
    if ((!z && n > 0) || !mu || n < 0 || !(alpha > 0.0)) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if (!ws || ws_bytes < workspace_bytes(d, K)) return SCC_ERR_WORKSPACE;
    return SCC_OK;
}

int dec_assign_f64(const double* z, int64_t n, int d, const double* mu, int K, double alpha, int round_decimals, double* q,
                   int32_t* labels, const int32_t* labels_prev, double* stats, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = check_f64(z, n, d, mu, K, alpha, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (!stats || (round_decimals != 0 && round_decimals != 5)) return SCC_ERR_INVALID;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K + 1), st)); return SCC_OK; }
    dec_assign_f64_kernel<<<f64_grid(n), kF64Threads, 0, st>>>(
        z, n, d, mu, K, alpha, round_decimals == 5, q, labels, labels_prev, stats,
        reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader), reinterpret_cast<unsigned int*>(ws));
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

int dec_target_f64(const double* q, int64_t n, int K, double* f, int have_f, int round_decimals, double* p, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
    if (!q || !f || !p || n < 0 || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if (round_decimals != 0 && round_decimals != 5) return SCC_ERR_INVALID;
    if (n == 0) return SCC_OK;
    if (!have_f) {
        if (!ws || ws_bytes < workspace_bytes(4, K)) return SCC_ERR_WORKSPACE;
        colsum_f64_kernel<<<f64_grid(n), kF64Threads, 0, st>>>(
            q, n, K, f, reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader),
            reinterpret_cast<unsigned int*>(ws));
        SCC_CUDA(cudaGetLastError());
    }
    dec_target_f64_kernel<<<f64_grid(n), kF64Threads, 0, st>>>(q, n, K, f, round_decimals == 5, p);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

int dec_grad_f64(const double* z, int64_t n, int d, const double* mu, int K, double alpha, const double* p,
                 const double* f_cols, int round_decimals, const double* grad_q, double scale, double* p_out, double* dz,
                 double* stats, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = check_f64(z, n, d, mu, K, alpha, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (!stats || (round_decimals != 0 && round_decimals != 5)) return SCC_ERR_INVALID;
    const int given = (p ? 1 : 0) + (f_cols ? 1 : 0) + (grad_q ? 1 : 0);
    if (given != 1 || (p_out && !f_cols)) return SCC_ERR_INVALID;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K * d + 2), st)); return SCC_OK; }
    const int mode = p ? 0 : (f_cols ? 1 : 2);
    const size_t smem = sizeof(double) * ((size_t)K * d + (kF64Threads / 32 + 1) * ((size_t)K * d + 2) + 2 * kF64Threads);
    SCC_CUDA(cudaFuncSetAttribute(dec_grad_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dec_grad_f64_kernel<<<f64_grid(n), kF64Threads, smem, st>>>(
        z, n, d, mu, K, alpha, mode, p, f_cols, round_decimals == 5, grad_q, scale, p_out, dz, stats,
        reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader), reinterpret_cast<unsigned int*>(ws));
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

}  // namespace scc
