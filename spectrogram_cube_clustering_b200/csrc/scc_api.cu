// scc_api.cu — the extern "C" boundary declared in include/scc_b200.h.
// Thin argument forwarding only; all validation lives next to the launchers.
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>

#include "scc_launch.h"

namespace scc {

static thread_local char g_cuda_error[512] = "";
unsigned long long* g_timeline = nullptr;

void set_cuda_error(cudaError_t e, const char* what, int line) {
    snprintf(g_cuda_error, sizeof(g_cuda_error), "%s (%s) at %s [line %d]", cudaGetErrorName(e),
             cudaGetErrorString(e), what, line);
}

namespace {
struct GridKey {
    const void* fn; int dev; size_t smem;
    bool operator<(const GridKey& o) const {
        if (fn != o.fn) return fn < o.fn;
        if (dev != o.dev) return dev < o.dev;
        return smem < o.smem;
    }
};
std::mutex g_grid_mu;
std::map<GridKey, int> g_grid_cache;
}  // namespace

int persistent_grid(const void* kernel, int threads, size_t smem, int max_ctas_per_sm) {
    int dev = 0;
    SCC_CUDA(cudaGetDevice(&dev));
    const GridKey key{kernel, dev, smem};
    {
        std::lock_guard<std::mutex> lk(g_grid_mu);
        auto it = g_grid_cache.find(key);
        if (it != g_grid_cache.end()) return it->second;
    }
    int sms = 0, occ = 0;
    SCC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SCC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SCC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    if (occ < 1) return SCC_ERR_UNSUPPORTED;
    if (occ > max_ctas_per_sm) occ = max_ctas_per_sm;
    const int grid = sms * occ;
    std::lock_guard<std::mutex> lk(g_grid_mu);
    g_grid_cache[key] = grid;
    return grid;
}

size_t workspace_bytes(int d, int K) {
    if (d < 1 || d > SCC_MAX_D || K < 1 || K > SCC_MAX_K) return 0;
    // per CTA: the tail's slot (grid_publish, even stride; MODE_KMEANS appends K counts) and — one-kernel step —
    // room for a pass-1 slot behind all tail slots and a K + 2 vector behind those (the one-kernel step's f barrier
    // now lives in the header's counted accumulators, see CountedFix; the space is kept for ABI stability)
    const size_t dec = (size_t)kMaxDecGrid * (size_t)(((K * d + 2 + K + 1) & ~1) + ((K + 2) & ~1)) + (size_t)(K + 2);
    const size_t gmm = (size_t)kMaxGmmGrid * (size_t)SCC_GMM_STAT_DOUBLES(K, d);
    return kWorkspaceHeader + sizeof(double) * (dec > gmm ? dec : gmm) + (size_t)(64 << 10);   // + staged GMM params
}

}  // namespace scc

extern "C" {

int scc_abi_version(void) { return SCC_ABI_VERSION; }

const char* scc_status_string(int status) {
    switch (status) {
        case SCC_OK: return "ok";
        case SCC_ERR_INVALID: return "invalid argument";
        case SCC_ERR_UNSUPPORTED: return "unsupported (d, K): no kernel instantiation";
        case SCC_ERR_MISALIGNED: return "misaligned pointer (16-byte alignment required)";
        case SCC_ERR_WORKSPACE: return "workspace missing or too small";
        case SCC_ERR_CUDA: return "CUDA runtime error";
        default: return "unknown status";
    }
}

const char* scc_last_cuda_error(void) { return scc::g_cuda_error; }

int scc_debug_set_timeline(void* device_buffer) {
#ifdef SCC_TIMELINE
    scc::g_timeline = reinterpret_cast<unsigned long long*>(device_buffer);
    return SCC_OK;
#else
    (void)device_buffer;
    return SCC_ERR_UNSUPPORTED;        // production build: the stamps are compiled out
#endif
}

int scc_supported(int d, int K) { return scc::dec_supported(d, K) ? 1 : 0; }
int scc_gmm_supported(int d, int K) { return scc::gmm_supported(d, K) ? 1 : 0; }

size_t scc_workspace_bytes(int d, int K) { return scc::workspace_bytes(d, K); }

int scc_workspace_init(void* workspace, size_t bytes, scc_stream_t stream) {
    if (!workspace || bytes < scc::kWorkspaceHeader) return SCC_ERR_WORKSPACE;
    SCC_CUDA(cudaMemsetAsync(workspace, 0, scc::kWorkspaceHeader, (cudaStream_t)stream));
    return SCC_OK;
}

int scc_dec_assign(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                   float* q, int32_t* labels, const int32_t* labels_prev, double* stats, void* workspace,
                   size_t workspace_bytes, scc_stream_t stream) {
    return scc::dec_assign(z, n, d, mu, K, alpha, round_decimals, q, labels, labels_prev, stats, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}

int scc_dec_target(const float* q, int64_t n, int K, const double* f, int round_decimals, float* p,
                   scc_stream_t stream) {
    return scc::dec_target(q, n, K, const_cast<double*>(f), round_decimals, p, (cudaStream_t)stream);
}

int scc_colsum(const float* q, int64_t n, int K, double* f, void* workspace, size_t workspace_bytes,
               scc_stream_t stream) {
    return scc::colsum(q, n, K, f, workspace, workspace_bytes, (cudaStream_t)stream);
}

int scc_dec_kl_grad(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* p,
                    const double* f_cols, int round_decimals, float scale, float* dz, double* stats,
                    void* workspace, size_t workspace_bytes, scc_stream_t stream) {
    return scc::dec_kl_grad(z, n, d, mu, K, alpha, p, f_cols, round_decimals, scale, dz, stats, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

int scc_dec_backward(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* grad_q,
                     float* dz, double* stats, void* workspace, size_t workspace_bytes, scc_stream_t stream) {
    return scc::dec_backward(z, n, d, mu, K, alpha, grad_q, dz, stats, workspace, workspace_bytes,
                             (cudaStream_t)stream);
}

int scc_kmeans_step(const float* z, int64_t n, int d, const float* centers, int K, int32_t* labels, float* mindist,
                    double* stats, void* workspace, size_t workspace_bytes, scc_stream_t stream) {
    return scc::kmeans_step(z, n, d, centers, K, labels, mindist, stats, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}

size_t scc_kmeans_batch_workspace_bytes(int d, int K, int restarts) {
    return scc::kmeans_batch_workspace_bytes(d, K, restarts);
}

int scc_kmeans_batch_step(const float* z, int64_t n, int d, const float* centers, int K, int restarts,
                          const unsigned char* done, int32_t* labels, float* mindist, double* stats, void* workspace,
                          size_t workspace_bytes, scc_stream_t stream) {
    return scc::kmeans_batch_step(z, n, d, centers, K, restarts, done, labels, mindist, stats, workspace,
                                  workspace_bytes, (cudaStream_t)stream);
}

int scc_kmeans_batch_update(float* centers, const double* stats, int d, int K, int restarts, double shift_tol,
                            unsigned char* done, int32_t* n_iter, double* inertia, scc_stream_t stream) {
    return scc::kmeans_batch_update(centers, stats, d, K, restarts, shift_tol, done, n_iter, inertia,
                                    (cudaStream_t)stream);
}

int scc_dec_distances(const float* z, int64_t n, int d, const float* mu, int K, float p, float* out,
                      scc_stream_t stream) {
    return scc::dec_distances(z, n, d, mu, K, p, out, (cudaStream_t)stream);
}

int scc_dec_assign_f64(const double* z, int64_t n, int d, const double* mu, int K, double alpha, int round_decimals,
                       double* q, int32_t* labels, const int32_t* labels_prev, double* stats, void* workspace,
                       size_t workspace_bytes, scc_stream_t stream) {
    return scc::dec_assign_f64(z, n, d, mu, K, alpha, round_decimals, q, labels, labels_prev, stats, workspace,
                               workspace_bytes, (cudaStream_t)stream);
}

int scc_dec_target_f64(const double* q, int64_t n, int K, double* f, int have_f, int round_decimals, double* p,
                       void* workspace, size_t workspace_bytes, scc_stream_t stream) {
    return scc::dec_target_f64(q, n, K, f, have_f, round_decimals, p, workspace, workspace_bytes, (cudaStream_t)stream);
}

int scc_dec_grad_f64(const double* z, int64_t n, int d, const double* mu, int K, double alpha, const double* p,
                     const double* f_cols, int round_decimals, const double* grad_q, double scale, double* p_out,
                     double* dz, double* stats, void* workspace, size_t workspace_bytes, scc_stream_t stream) {
    return scc::dec_grad_f64(z, n, d, mu, K, alpha, p, f_cols, round_decimals, grad_q, scale, p_out, dz, stats, workspace,
                             workspace_bytes, (cudaStream_t)stream);
}

static const scc::ExchangeDesc* as_desc(const scc_exchange* e, scc::ExchangeDesc* tmp) {
    if (!e || !e->windows) return nullptr;
    tmp->windows = e->windows; tmp->rank = e->rank; tmp->world = e->world; tmp->max_len = e->max_len;
    return tmp;
}

int scc_dec_assign_ex(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                      float* q, int32_t* labels, const int32_t* labels_prev, double* stats, void* workspace,
                      size_t workspace_bytes, const scc_exchange* push, scc_stream_t stream) {
    scc::ExchangeDesc t;
    return scc::dec_assign(z, n, d, mu, K, alpha, round_decimals, q, labels, labels_prev, stats, workspace,
                           workspace_bytes, (cudaStream_t)stream, as_desc(push, &t));
}

int scc_dec_target_ex(const float* q, int64_t n, int K, double* f, int round_decimals, float* p,
                      const scc_exchange* pull, scc_stream_t stream) {
    scc::ExchangeDesc t;
    return scc::dec_target(q, n, K, f, round_decimals, p, (cudaStream_t)stream, as_desc(pull, &t));
}

int scc_dec_kl_grad_ex(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* p,
                       const double* f_cols, int round_decimals, float scale, float* dz, double* stats,
                       void* workspace, size_t workspace_bytes, const scc_exchange* pull_f,
                       const scc_exchange* push, scc_stream_t stream) {
    scc::ExchangeDesc t1, t2;
    return scc::dec_kl_grad(z, n, d, mu, K, alpha, p, f_cols, round_decimals, scale, dz, stats, workspace,
                            workspace_bytes, (cudaStream_t)stream, as_desc(pull_f, &t1), as_desc(push, &t2));
}

int scc_dec_target_kl_grad(const float* z, int64_t n, int d, const float* mu, int K, float alpha,
                           const double* f_cols, int round_decimals, float scale, float* p_out, float* dz,
                           double* stats, void* workspace, size_t workspace_bytes, const scc_exchange* pull_f,
                           const scc_exchange* push, scc_stream_t stream) {
    scc::ExchangeDesc t1, t2;
    return scc::dec_kl_grad(z, n, d, mu, K, alpha, nullptr, f_cols, round_decimals, scale, dz, stats, workspace,
                            workspace_bytes, (cudaStream_t)stream, as_desc(pull_f, &t1), as_desc(push, &t2), p_out);
}

int scc_dec_assign_u(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                     float* q, int32_t* labels, const int32_t* labels_prev, float* u_out, double* stats, void* workspace,
                     size_t workspace_bytes, const scc_exchange* push, scc_stream_t stream) {
    scc::ExchangeDesc t;
    return scc::dec_assign(z, n, d, mu, K, alpha, round_decimals, q, labels, labels_prev, stats, workspace,
                           workspace_bytes, (cudaStream_t)stream, as_desc(push, &t), u_out);
}

int scc_dec_target_kl_grad_u(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* u_in,
                             const double* f_cols, int round_decimals, float scale, float* p_out, float* dz,
                             double* stats, void* workspace, size_t workspace_bytes, const scc_exchange* pull_f,
                             const scc_exchange* push, scc_stream_t stream) {
    if (!u_in) return SCC_ERR_INVALID;
    scc::ExchangeDesc t1, t2;
    return scc::dec_kl_grad(z, n, d, mu, K, alpha, nullptr, f_cols, round_decimals, scale, dz, stats, workspace,
                            workspace_bytes, (cudaStream_t)stream, as_desc(pull_f, &t1), as_desc(push, &t2), p_out, u_in);
}

int scc_dec_step(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                 float scale, float* q, int32_t* labels, const int32_t* labels_prev, double* f_stats, float* p_out,
                 float* dz, double* stats, void* workspace, size_t workspace_bytes, scc_stream_t stream) {
    return scc::dec_step(z, n, d, mu, K, alpha, round_decimals, scale, q, labels, labels_prev, f_stats, p_out, dz,
                         stats, workspace, workspace_bytes, (cudaStream_t)stream);
}

int scc_dec_step_ex(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                    float scale, float* q, int32_t* labels, const int32_t* labels_prev, double* f_stats, float* p_out,
                    float* dz, double* stats, void* workspace, size_t workspace_bytes, const scc_exchange* ex,
                    scc_stream_t stream) {
    scc::ExchangeDesc t;
    return scc::dec_step(z, n, d, mu, K, alpha, round_decimals, scale, q, labels, labels_prev, f_stats, p_out, dz,
                         stats, workspace, workspace_bytes, (cudaStream_t)stream, as_desc(ex, &t));
}

int scc_peer_finish(double* out, int len, const scc_exchange* ex, scc_stream_t stream) {
    if (!ex || !ex->windows) return SCC_ERR_INVALID;
    return scc::peer_finish(out, len, ex->windows, ex->rank, ex->world, ex->max_len, (cudaStream_t)stream);
}

size_t scc_peer_window_bytes(int max_len) { return scc::peer_window_bytes(max_len); }

int scc_peer_allreduce(const double* local, int len, double* out, void* const* peer_windows, int rank, int world,
                       int max_len, scc_stream_t stream) {
    return scc::peer_allreduce(local, len, out, peer_windows, rank, world, max_len, (cudaStream_t)stream);
}

int scc_gmm_em_step(const float* z, int64_t n, int d, int K, const float* params, double* stats, int32_t* labels,
                    float* resp, const double* ctrl, int mode, void* workspace,
                    size_t workspace_bytes, scc_stream_t stream) {
    return scc::gmm_em_step(z, n, d, K, params, stats, labels, resp, ctrl, mode, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

int scc_gmm_finalize(const double* stats, double n_total, int d, int K, double reg_covar, double nk_eps, double tol,
                     double* means, double* weights, double* covariances, double* prec_chol, float* params,
                     double* ctrl, scc_stream_t stream) {
    return scc::gmm_finalize(stats, n_total, d, K, reg_covar, nk_eps, tol, means, weights, covariances, prec_chol,
                             params, ctrl, (cudaStream_t)stream);
}

int scc_gmm_em_iteration(const float* z, int64_t n, int d, int K, float* params, double* stats, int mode,
                         double n_total, double reg_covar, double nk_eps, double tol, double* means, double* weights,
                         double* covariances, double* prec_chol, double* ctrl, void* workspace, size_t workspace_bytes,
                         const scc_exchange* exchange, scc_stream_t stream) {
    scc::ExchangeDesc t;
    return scc::gmm_em_iteration(z, n, d, K, params, stats, mode, n_total, reg_covar, nk_eps, tol, means, weights,
                                 covariances, prec_chol, ctrl, workspace, workspace_bytes, as_desc(exchange, &t),
                                 (cudaStream_t)stream);
}

int scc_gmm_pack_params(const double* weights, const double* means, const double* covariances, int d, int K,
                        double* prec_chol, float* params, double* ctrl, scc_stream_t stream) {
    return scc::gmm_pack_params(weights, means, covariances, d, K, prec_chol, params, ctrl, (cudaStream_t)stream);
}

}  // extern "C"
