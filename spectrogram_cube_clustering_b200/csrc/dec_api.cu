// dec_api.cu — DEC entry points: argument validation, per-dimension dispatch, and the two
// dimension-independent kernels (target distribution, column sums).
#include "dec_kernels.cuh"

namespace scc {

#define SCC_DECL_DIM(D_, X)                                             \
    int dec_assign_dim##D_(const DecArgs& a, cudaStream_t st);          \
    int dec_grad_dim##D_(const DecArgs& a, int mode, cudaStream_t st);
SCC_FOR_EACH_DIM(SCC_DECL_DIM, 0)
#undef SCC_DECL_DIM

// ---------------------------------------------------------------------------
// dec_target: p = normalise_rows(q^2 / f)   (models.py:1320-1322)
// LPR lanes cooperate on one row (K = 4*LPR) so that global accesses are
// 128-bit and fully coalesced; LPR = 0 is the scalar thread-per-row fallback.
// ---------------------------------------------------------------------------
// kept out of line so that the exchange prologue does not inflate the streaming loop's registers
static __device__ __noinline__ void target_pull_f(const PeerCtx& pull, int K, double* f, float* inv_f) {
    __shared__ double f_pull[SCC_MAX_K + 1];
    peer_pull(pull, f_pull, K + 1);
    if ((int)threadIdx.x < K) inv_f[threadIdx.x] = (float)(1.0 / f_pull[threadIdx.x]);
    if (blockIdx.x == 0 && (int)threadIdx.x <= K) f[threadIdx.x] = f_pull[threadIdx.x];   // publish the sums
}

template <int LPR>
__global__ void __launch_bounds__(256, 8)          // <= 32 registers: 8 CTAs per SM keep enough loads in flight to stream at HBM rate
dec_target_kernel(const float* __restrict__ q, int64_t n, int K, double* __restrict__ f,
                  int round5, float* __restrict__ p, PeerCtx pull) {
    __shared__ float inv_f[SCC_MAX_K];
    pdl_wait();                         // q and f are predecessor outputs
    if (pull.windows) {                 // f comes from the exchange pushed by the preceding assign kernel
        target_pull_f(pull, K, f, inv_f);
    } else if (threadIdx.x < K) {
        inv_f[threadIdx.x] = (float)(1.0 / f[threadIdx.x]);
    }
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
// colsum: f_j = sum_i q_ij for a caller-supplied q (models.py:1320)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecThreads)
colsum_kernel(const float* __restrict__ q, int64_t n, int K, double* stats, double* partials, unsigned int* counter) {
    constexpr int KP = SCC_MAX_K;
    __shared__ double scratch[reduce_scratch(KP)];
    __shared__ double cta_stats[KP];
    float acc[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) acc[j] = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v[KP];
        load_krow<KP, false>(q + i * K, K, v);
#pragma unroll
        for (int j = 0; j < KP; ++j) acc[j] += v[j];
    }
    cta_reduce<KP, kDecThreads>(acc, scratch, cta_stats);
    grid_publish<kDecThreads>(cta_stats, K, partials, counter, stats, scratch);
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


int dec_target(const float* q, int64_t n, int K, double* f, int round_decimals, float* p, cudaStream_t st,
               const ExchangeDesc* pull) {
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
    PeerCtx px{nullptr, 0, 1, 0};
    if (pull && pull->windows) px = PeerCtx{reinterpret_cast<unsigned char* const*>(pull->windows), pull->rank, pull->world, pull->max_len};
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // PDL, see scc_common.cuh
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (vec && K == 4) SCC_CUDA(cudaLaunchKernelEx(&cfg, dec_target_kernel<1>, q, n, K, f, r5, p, px));
    else if (vec && K == 8) SCC_CUDA(cudaLaunchKernelEx(&cfg, dec_target_kernel<2>, q, n, K, f, r5, p, px));
    else if (vec && K == 16) SCC_CUDA(cudaLaunchKernelEx(&cfg, dec_target_kernel<4>, q, n, K, f, r5, p, px));
    else SCC_CUDA(cudaLaunchKernelEx(&cfg, dec_target_kernel<0>, q, n, K, f, r5, p, px));
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}


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
    if (n > (int64_t)kDecTile * 1000000000LL) return SCC_ERR_INVALID;
    if (!dec_supported(d, K)) return SCC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(z) & 15u) != 0) return SCC_ERR_MISALIGNED;
    if (!ws || ws_bytes < workspace_bytes(d, K)) return SCC_ERR_WORKSPACE;
    return SCC_OK;
}

static void fill_exchange(DecArgs& a, const ExchangeDesc* push, int do_push, const ExchangeDesc* pull_f) {
    const ExchangeDesc* e = (push && push->windows) ? push : ((pull_f && pull_f->windows) ? pull_f : nullptr);
    if (!e) return;
    a.ex_windows = reinterpret_cast<unsigned char* const*>(e->windows);
    a.ex_rank = e->rank; a.ex_world = e->world; a.ex_max_len = e->max_len;
    a.ex_push = (push && push->windows) ? do_push : 0;
    a.ex_pull_f = (pull_f && pull_f->windows) ? 1 : 0;
}

static void fill_reduction(DecArgs& a, void* ws) {
    a.counter = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kWorkspaceHeader);
    a.timeline = g_timeline;
}

int dec_assign(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
               float* q, int32_t* labels, const int32_t* labels_prev, double* stats,
               void* ws, size_t ws_bytes, cudaStream_t st, const ExchangeDesc* push, float* u_out) {
    int rc = check_common(z, n, d, mu, K, alpha, stats, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (round_decimals != 0 && round_decimals != 5) return SCC_ERR_INVALID;
    if (q && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(q) & 15u)) return SCC_ERR_MISALIGNED;
    if (u_out && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(u_out) & 15u)) return SCC_ERR_MISALIGNED;
    if (n == 0) {
        SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K + 1), st));
        if (push && push->windows) return peer_push_only(stats, K + 1, push->windows, push->rank, push->world, push->max_len, st);
        return SCC_OK;
    }
    DecArgs a{};
    a.z = z; a.n = n; a.mu = mu; a.K = K; a.alpha = alpha; a.round5 = round_decimals == 5;
    a.q = q; a.labels = labels; a.labels_prev = labels_prev; a.stats = stats; a.u_out = u_out;
    fill_reduction(a, ws);
    fill_exchange(a, push, /*push=*/1, nullptr);
#define SCC_CASE(D_, X) if (d == D_) return dec_assign_dim##D_(a, st);
    SCC_FOR_EACH_DIM(SCC_CASE, 0)
#undef SCC_CASE
    return SCC_ERR_UNSUPPORTED;
}

static int dec_grad_dispatch(const DecArgs& a, int d, int mode, cudaStream_t st) {
#define SCC_CASE(D_, X) if (d == D_) return dec_grad_dim##D_(a, mode, st);
    SCC_FOR_EACH_DIM(SCC_CASE, 0)
#undef SCC_CASE
    return SCC_ERR_UNSUPPORTED;
}

int dec_kl_grad(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* p,
                const double* f_cols, int round_decimals, float scale, float* dz, double* stats,
                void* ws, size_t ws_bytes, cudaStream_t st, const ExchangeDesc* pull_f, const ExchangeDesc* push,
                float* p_out, const float* u_in) {
    int rc = check_common(z, n, d, mu, K, alpha, stats, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (!p && !f_cols && !(pull_f && pull_f->windows)) return SCC_ERR_INVALID;
    if (u_in && p) return SCC_ERR_INVALID;             // the hand-off replaces the distance loop of the fused mode
    if (u_in && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(u_in) & 15u)) return SCC_ERR_MISALIGNED;
    if (p && p_out) return SCC_ERR_INVALID;            // the target is either given or produced, not both
    if (p_out && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(p_out) & 15u)) return SCC_ERR_MISALIGNED;
    if (round_decimals != 0 && round_decimals != 5) return SCC_ERR_INVALID;
    if (p && (K % 4 == 0) && (reinterpret_cast<uintptr_t>(p) & 15u)) return SCC_ERR_MISALIGNED;
    if (dz && (reinterpret_cast<uintptr_t>(dz) & 15u)) return SCC_ERR_MISALIGNED;
    if (n == 0) {
        SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K * d + 2), st));
        if (push && push->windows)      // an empty shard still takes part in the (in-kernel) all-reduce
            return peer_allreduce(stats, K * d + 2, stats, push->windows, push->rank, push->world, push->max_len, st);
        return SCC_OK;
    }
    DecArgs a{};
    a.z = z; a.n = n; a.mu = mu; a.K = K; a.alpha = alpha; a.round5 = round_decimals == 5;
    a.p = p; a.p_out = p_out; a.f_cols = f_cols; a.scale = scale; a.dz = dz; a.stats = stats; a.u_in = u_in;
    fill_reduction(a, ws);
    fill_exchange(a, push, /*push + collect in the kernel's tail=*/2, p ? nullptr : pull_f);
    return dec_grad_dispatch(a, d, p ? MODE_KL : (u_in ? MODE_KLU : MODE_KLF), st);
}

int dec_step(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals, float scale,
             float* q, int32_t* labels, const int32_t* labels_prev, double* f_stats, float* p_out, float* dz,
             double* stats, void* ws, size_t ws_bytes, cudaStream_t st, const ExchangeDesc* ex) {
    int rc = check_common(z, n, d, mu, K, alpha, stats, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (!f_stats) return SCC_ERR_INVALID;
    if (round_decimals != 0 && round_decimals != 5) return SCC_ERR_INVALID;
    if ((K % 4 == 0) && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(p_out)) & 15u))
        return SCC_ERR_MISALIGNED;
    if (dz && (reinterpret_cast<uintptr_t>(dz) & 15u)) return SCC_ERR_MISALIGNED;
    const bool multi = ex && ex->windows;
    if (n == 0 && !multi) {         // (an empty shard of a multi-GPU step still runs: it takes part in the exchanges)
        SCC_CUDA(cudaMemsetAsync(f_stats, 0, sizeof(double) * (K + 1), st));
        SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K * d + 2), st));
        return SCC_OK;
    }
    DecArgs a{};
    a.z = z; a.n = n; a.mu = mu; a.K = K; a.alpha = alpha; a.round5 = round_decimals == 5;
    a.q = q; a.labels = labels; a.labels_prev = labels_prev; a.f_out = f_stats;
    a.p_out = p_out; a.scale = scale; a.dz = dz; a.stats = stats;
    fill_reduction(a, ws);
    fill_exchange(a, ex, /*all-reduce the final statistics in the kernel's tail=*/2, nullptr);
    return dec_grad_dispatch(a, d, MODE_STEP, st);
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
    return dec_grad_dispatch(a, d, MODE_GENERIC, st);
}

// ---------------------------------------------------------------------------
// Batched Lloyd iterations: R restarts of KMeans(n_init=R) advance together — one launch scans z against
// R x K centres (grid.y = restart), a second tiny launch moves the centres and decides convergence per
// restart on the device (sklearn: total squared centre shift <= tol * mean feature variance).
// ---------------------------------------------------------------------------
static size_t kmeans_batch_header(int R) { return ((size_t)R * 2 * sizeof(unsigned int) + 255) & ~(size_t)255; }

size_t kmeans_batch_workspace_bytes(int d, int K, int R) {
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K || R < 1) return 0;
    const size_t sp = (size_t)((K * d + 2 + K + 1) & ~1);
    return kmeans_batch_header(R) + sizeof(double) * (size_t)R * kBatchGridX * sp;
}

int kmeans_batch_step(const float* z, int64_t n, int d, const float* centers, int K, int R, const unsigned char* done,
                      int32_t* labels, float* mindist, double* stats, void* ws, size_t ws_bytes, cudaStream_t st) {
    if ((!z && n > 0) || !centers || !stats || n < 0 || R < 1 || R > 65535) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if (!dec_supported(d, K)) return SCC_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(z) & 15u) != 0) return SCC_ERR_MISALIGNED;
    if (!ws || ws_bytes < kmeans_batch_workspace_bytes(d, K, R)) return SCC_ERR_WORKSPACE;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (size_t)R * (K * d + 2 + K), st)); return SCC_OK; }
    DecArgs a{};
    a.z = z; a.n = n; a.mu = centers; a.K = K; a.alpha = 1.0f; a.scale = 1.f;
    a.labels = labels; a.mindist = mindist; a.stats = stats;
    a.batch = R; a.batch_done = done;
    a.counter = reinterpret_cast<unsigned int*>(ws);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + kmeans_batch_header(R));
    a.timeline = nullptr;
    return dec_grad_dispatch(a, d, MODE_KMEANS, st);
}

__global__ void __launch_bounds__(128)
kmeans_batch_update_kernel(float* __restrict__ centers, const double* __restrict__ stats, int d, int K, double thresh,
                           unsigned char* __restrict__ done, int32_t* __restrict__ n_iter, double* __restrict__ inertia) {
    const int r = blockIdx.x;
    if (done[r]) return;
    __shared__ double red[128];
    const int S = K * d + 2 + K;
    const double* st = stats + (size_t)r * S;
    float* c = centers + (size_t)r * K * d;
    double sh = 0.0;
    for (int o = threadIdx.x; o < K * d; o += blockDim.x) {
        const double cnt = st[2 + K * d + o / d];
        const double step = cnt > 0.0 ? st[2 + o] / cnt : 0.0;      // an empty cluster keeps its centre
        c[o] = (float)((double)c[o] + step);
        sh += step * step;
    }
    red[threadIdx.x] = sh;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int t = 0; t < (int)blockDim.x; ++t) tot += red[t];
        if (inertia) inertia[r] = st[0];                             // inertia of the centres the scan used
        if (n_iter) n_iter[r] += 1;
        if (tot <= thresh) done[r] = 1;
    }
}

int kmeans_batch_update(float* centers, const double* stats, int d, int K, int R, double thresh, unsigned char* done,
                        int32_t* n_iter, double* inertia, cudaStream_t st) {
    if (!centers || !stats || !done || R < 1) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    kmeans_batch_update_kernel<<<R, 128, 0, st>>>(centers, stats, d, K, thresh, done, n_iter, inertia);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

// ---------------------------------------------------------------------------
// dec_distances: D_ij = (sum_c |z_ic - mu_jc|^p)^(1/p) for every point and centroid — the scan behind
// utils.fractional_distance / distance_matrix (utils.py:866-869, 635-643) and, with p = 2 squared and summed,
// utils.measure_class_inertia (utils.py:1024-1029).  HBM-bound: reads z once, writes [n, K].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dec_distances_kernel(const float* __restrict__ z, int64_t n, int d, const float* __restrict__ mu, int K, float p,
                     float* __restrict__ out) {
    extern __shared__ float dist_smem[];
    float* mu_s = dist_smem;                       // [K*d]
    float* rows = dist_smem + K * d;               // [256][d+1] (odd-ish stride: conflict-free row reads)
    const int ld = d + 1;
    for (int i = threadIdx.x; i < K * d; i += 256) mu_s[i] = mu[i];
    const float inv_p = 1.f / p;
    const bool p2 = (p == 2.f), p1 = (p == 1.f);
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += (int64_t)gridDim.x * 256) {
        const int np = (int)((n - base < 256) ? (n - base) : 256);
        __syncthreads();
        for (int f = threadIdx.x; f < np * d; f += 256) {
            const int row = f / d, c = f - row * d;
            rows[row * ld + c] = ldg_stream(z + base * d + f);
        }
        __syncthreads();
        if ((int)threadIdx.x < np) {
            const float* x = rows + threadIdx.x * ld;
            for (int j = 0; j < K; ++j) {
                float acc = 0.f;
                for (int c = 0; c < d; ++c) {
                    const float df = fabsf(x[c] - mu_s[j * d + c]);
                    acc += p2 ? df * df : (p1 ? df : powf(df, p));
                }
                out[(base + threadIdx.x) * K + j] = p2 ? sqrtf(acc) : (p1 ? acc : powf(acc, inv_p));
            }
        }
    }
}

int dec_distances(const float* z, int64_t n, int d, const float* mu, int K, float p, float* out, cudaStream_t st) {
    if ((!z && n > 0) || !mu || (!out && n > 0) || n < 0 || !(p > 0.f)) return SCC_ERR_INVALID;
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return SCC_ERR_INVALID;
    if (n == 0) return SCC_OK;
    int dev = 0, sms = 0;
    SCC_CUDA(cudaGetDevice(&dev));
    SCC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int64_t grid = (n + 255) / 256;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;
    const size_t smem = sizeof(float) * ((size_t)K * d + 256 * (size_t)(d + 1));
    dec_distances_kernel<<<(unsigned)grid, 256, smem, st>>>(z, n, d, mu, K, p, out);
    SCC_CUDA(cudaGetLastError());
    return SCC_OK;
}

int kmeans_step(const float* z, int64_t n, int d, const float* centers, int K, int32_t* labels, float* mindist,
                double* stats, void* ws, size_t ws_bytes, cudaStream_t st) {
    int rc = check_common(z, n, d, centers, K, 1.0f, stats, ws, ws_bytes);
    if (rc != SCC_OK) return rc;
    if (n == 0) { SCC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * (K * d + 2 + K), st)); return SCC_OK; }
    DecArgs a{};
    a.z = z; a.n = n; a.mu = centers; a.K = K; a.alpha = 1.0f; a.scale = 1.f;
    a.labels = labels; a.mindist = mindist; a.stats = stats;
    fill_reduction(a, ws);
    return dec_grad_dispatch(a, d, MODE_KMEANS, st);
}

}  // namespace scc
