// scc_launch.h — host-side declarations shared by the kernel translation units
// and the C-ABI glue (scc_api.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "scc_b200.h"

namespace scc {

// Latent dimensions with kernel instantiations.  Other d in [1,32] are served
// by the Python host through zero-padding of z and mu (exact for the DEC path).
#define SCC_FOR_EACH_DIM(M, A) M(4, A) M(8, A) M(9, A) M(10, A) M(12, A) M(16, A) M(20, A) M(24, A) M(32, A)

constexpr int kMaxCtasPerSm = 8;
constexpr int kMaxDecGrid = 2048;      // persistent DEC grids never exceed this many CTAs
constexpr int kMaxGmmGrid = 512;
constexpr size_t kWorkspaceHeader = 256;   // counters live in front of the partial slots

// MODE_KL: target p streamed from memory;  MODE_KLF: p rebuilt in registers from the column sums (fused mode)
// MODE_STEP: MODE_KLF preceded, in the same (cooperative) kernel, by the assign pass and a grid-wide all-reduce of f
// MODE_KLU: MODE_KLF with the Student's-t u_ij streamed from memory (written by the assign pass) instead of recomputed
enum { MODE_KL = 0, MODE_GENERIC = 1, MODE_KMEANS = 2, MODE_KLF = 3, MODE_STEP = 4, MODE_KLU = 5 };

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
    float* mindist;            // MODE_KMEANS: squared distance to the nearest centre, or NULL
    float* u_out;              // assign: [n, K] u_ij = 1 / (1 + d_ij / alpha) for a following MODE_KLU pass, or NULL
    const float* u_in;         // MODE_KLU
    // grad
    const float* p;
    float* p_out;              // fused mode (p == NULL): also write the rebuilt target rows here, or NULL
    const double* f_cols;
    double* f_out;             // MODE_STEP: [K+1] column sums + label-change count of the assign pass
    const float* grad_q;
    float scale;
    float* dz;
    // reduction
    double* stats;
    double* partials;
    unsigned int* counter;
    // optional fused multi-GPU exchange (windows == nullptr: none)
    unsigned char* const* ex_windows;
    int ex_rank, ex_world, ex_max_len;
    int ex_push;               // 1: last CTA pushes the reduced stats to every rank's window; 2: and collects the
                               // world's sum into `stats` itself (all-reduce complete when the kernel ends)
    int ex_pull_f;             // f_cols are pulled from the exchange pushed by the preceding assign kernel
    // batched Lloyd step (MODE_KMEANS): `batch` > 0 restarts over the same z, see batch_view()
    int batch;
    const unsigned char* batch_done;
    unsigned long long* timeline;   // profiling builds (-DSCC_TIMELINE): [grid][8] %globaltimer stamps, or NULL
};

// Device buffer the next DEC launches stamp their phase times into (profiling builds only).
extern unsigned long long* g_timeline;

struct ExchangeDesc {          // host-side mirror of scc_exchange
    void* const* windows;
    int rank, world, max_len;
};

void set_cuda_error(cudaError_t e, const char* what, int line);

// Cached per-kernel launch configuration: opts the kernel into `smem` bytes of dynamic shared
// memory once and returns the persistent-grid size (SMs x resident CTAs per SM, capped).
// Returns a negative scc_status on failure.
int persistent_grid(const void* kernel, int threads, size_t smem, int max_ctas_per_sm);
size_t workspace_bytes(int d, int K);
bool dec_supported(int d, int K);
bool gmm_supported(int d, int K);

#define SCC_CUDA(expr)                                                   \
    do {                                                                 \
        cudaError_t scc_e_ = (expr);                                     \
        if (scc_e_ != cudaSuccess) {                                     \
            ::scc::set_cuda_error(scc_e_, #expr, __LINE__);              \
            return SCC_ERR_CUDA;                                         \
        }                                                                \
    } while (0)

int dec_assign(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
               float* q, int32_t* labels, const int32_t* labels_prev, double* stats,
               void* ws, size_t ws_bytes, cudaStream_t st, const ExchangeDesc* push = nullptr, float* u_out = nullptr);
int dec_target(const float* q, int64_t n, int K, double* f, int round_decimals, float* p, cudaStream_t st,
               const ExchangeDesc* pull = nullptr);
int peer_finish(double* out, int len, void* const* windows_dev, int rank, int world, int max_len, cudaStream_t st);
int peer_push_only(const double* local, int len, void* const* windows_dev, int rank, int world, int max_len,
                   cudaStream_t st);
int colsum(const float* q, int64_t n, int K, double* f, void* ws, size_t ws_bytes, cudaStream_t st);
int dec_kl_grad(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* p,
                const double* f_cols, int round_decimals, float scale, float* dz, double* stats,
                void* ws, size_t ws_bytes, cudaStream_t st, const ExchangeDesc* pull_f = nullptr,
                const ExchangeDesc* push = nullptr, float* p_out = nullptr, const float* u_in = nullptr);
int dec_step(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals, float scale,
             float* q, int32_t* labels, const int32_t* labels_prev, double* f_stats, float* p_out, float* dz,
             double* stats, void* ws, size_t ws_bytes, cudaStream_t st, const ExchangeDesc* ex = nullptr);
int dec_backward(const float* z, int64_t n, int d, const float* mu, int K, float alpha, const float* grad_q,
                 float* dz, double* stats, void* ws, size_t ws_bytes, cudaStream_t st);
int kmeans_step(const float* z, int64_t n, int d, const float* centers, int K, int32_t* labels, float* mindist,
                double* stats, void* ws, size_t ws_bytes, cudaStream_t st);
size_t kmeans_batch_workspace_bytes(int d, int K, int R);
int kmeans_batch_step(const float* z, int64_t n, int d, const float* centers, int K, int R, const unsigned char* done,
                      int32_t* labels, float* mindist, double* stats, void* ws, size_t ws_bytes, cudaStream_t st);
int kmeans_batch_update(float* centers, const double* stats, int d, int K, int R, double thresh, unsigned char* done,
                        int32_t* n_iter, double* inertia, cudaStream_t st);
int dec_distances(const float* z, int64_t n, int d, const float* mu, int K, float p, float* out, cudaStream_t st);

// float64 precision path (dec_f64.cu)
int dec_assign_f64(const double* z, int64_t n, int d, const double* mu, int K, double alpha, int round_decimals, double* q,
                   int32_t* labels, const int32_t* labels_prev, double* stats, void* ws, size_t ws_bytes, cudaStream_t st);
int dec_target_f64(const double* q, int64_t n, int K, double* f, int have_f, int round_decimals, double* p, void* ws,
                   size_t ws_bytes, cudaStream_t st);
int dec_grad_f64(const double* z, int64_t n, int d, const double* mu, int K, double alpha, const double* p,
                 const double* f_cols, int round_decimals, const double* grad_q, double scale, double* p_out, double* dz,
                 double* stats, void* ws, size_t ws_bytes, cudaStream_t st);

size_t peer_window_bytes(int max_len);
int peer_allreduce(const double* local, int len, double* out, void* const* windows_dev, int rank, int world,
                   int max_len, cudaStream_t st);

int gmm_em_step(const float* z, int64_t n, int d, int K, const float* params, double* stats,
                int32_t* labels, float* resp, const double* ctrl, int mode,
                void* ws, size_t ws_bytes, cudaStream_t st);
int gmm_finalize(const double* stats, double n_total, int d, int K, double reg_covar, double nk_eps, double tol,
                 double* means, double* weights, double* covariances, double* prec_chol, float* params,
                 double* ctrl, cudaStream_t st);
int gmm_em_iteration(const float* z, int64_t n, int d, int K, float* params, double* stats, int mode, double n_total,
                     double reg_covar, double nk_eps, double tol, double* means, double* weights, double* covariances,
                     double* prec_chol, double* ctrl, void* ws, size_t ws_bytes, const ExchangeDesc* ex, cudaStream_t st);
int gmm_pack_params(const double* weights, const double* means, const double* covariances, int d, int K,
                    double* prec_chol, float* params, double* ctrl, cudaStream_t st);

}  // namespace scc
