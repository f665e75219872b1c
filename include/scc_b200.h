/*
 * scc_b200.h — C ABI of the B200-native latent-space clustering hot path.
 *
 * Drop-in boundary for ONE path of Julia310/Spectrogram-Cube-Clustering: the DEC
 * clustering layer (soft assignment, target distribution, KL loss + gradients)
 * and the full-covariance GMM E-step / M-step that seeds the centroids.  The
 * reference is pure Python (torch + scikit-learn) and has no FFI of its own;
 * each entry point below names the reference code (file:line under
 * /root/reference, or `sklearn:` for scikit-learn 1.9.0) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer
 * would add.
 *
 * Conventions
 *  - Plain C: raw DEVICE pointers, sizes, scalars, a CUDA stream handle
 *    (cudaStream_t passed as void*; NULL = legacy default stream).  No torch types.
 *  - Every function returns SCC_OK (0) or a negative scc_status; nothing is
 *    thrown, nothing is printed.  Launches are asynchronous on `stream`.
 *  - Latent points z are row-major [n, d] float32, 16-byte aligned base pointer.
 *    Centroids / means are row-major [K, d].  d in [1,32], K in [1,16];
 *    scc_supported(d, K) says whether kernels are instantiated for (d, K).
 *  - Statistics that are sums over points come back as float64 in a caller
 *    buffer (`stats`), already reduced over the grid in a fixed order
 *    (deterministic for a given device), ready for a cross-GPU allreduce(sum).
 *  - `workspace` is a caller-owned device scratch buffer of at least
 *    scc_workspace_bytes(d, K) bytes, zeroed ONCE with scc_workspace_init and
 *    then reusable by every call on the same stream (calls that share a
 *    workspace must be stream-ordered).
 *  - There is no CPU fallback: on a machine without an sm_100 device every
 *    compute entry point returns SCC_ERR_CUDA.
 */
#ifndef SCC_B200_H_
#define SCC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCC_ABI_VERSION 3
#define SCC_MAX_D 32
#define SCC_MAX_K 16

typedef enum scc_status {
    SCC_OK = 0,
    SCC_ERR_INVALID = -1,      /* null pointer, n < 0, K or d out of range, bad flag   */
    SCC_ERR_UNSUPPORTED = -2,  /* (d, K) has no kernel instantiation                   */
    SCC_ERR_MISALIGNED = -3,   /* a pointer violates the documented alignment          */
    SCC_ERR_WORKSPACE = -4,    /* workspace NULL or smaller than scc_workspace_bytes   */
    SCC_ERR_CUDA = -5          /* a CUDA runtime call failed; see scc_last_cuda_error  */
} scc_status;

typedef void* scc_stream_t;    /* cudaStream_t */

int scc_abi_version(void);
const char* scc_status_string(int status);
const char* scc_last_cuda_error(void);          /* thread-local text of the last CUDA failure */
int scc_supported(int d, int K);                /* DEC kernels instantiated for (d, K)? 1 / 0 */
/* Profiling hook (no reference counterpart).  In the `make timeline` build of the library the DEC
 * kernels stamp %globaltimer at their phase boundaries into device_buffer[grid][8] (uint64);
 * pass NULL to stop.  The production library compiles the stamps out and returns SCC_ERR_UNSUPPORTED. */
int scc_debug_set_timeline(void* device_buffer);
int scc_gmm_supported(int d, int K);            /* GMM kernels instantiated for (d, K)? 1 / 0 */

size_t scc_workspace_bytes(int d, int K);
int scc_workspace_init(void* workspace, size_t bytes, scc_stream_t stream);

/* ------------------------------------------------------------------------- *
 * DEC stage
 * ------------------------------------------------------------------------- */

/* Length (in doubles) of the stats buffer each DEC call fills. */
#define SCC_DEC_ASSIGN_STATS(K) ((K) + 1)             /* f[K], n_label_changes            */
#define SCC_DEC_GRAD_STATS(K, d) ((K) * (d) + 2)       /* loss, sum_i s_i, dmu[K*d]        */

/*
 * Student's-t soft assignment  q_ij = t_ij / sum_j t_ij,
 * t_ij = (1 + ||z_i - mu_j||^2 / alpha)^-((alpha+1)/2)
 *   replaces Cluster/networks.py:279-288 (ClusteringLayer.forward),
 *   the argmax of Cluster/models.py:92, the np.round(q, 5) of models.py:94,
 *   the column sums f_j = sum_i q_ij of models.py:1320 and the label-change
 *   count of models.py:1098-1099 — one pass over z.
 *
 * round_decimals: 0 keeps q as computed; 5 reproduces np.round(q, 5) and the
 *   column sums are then taken over the ROUNDED q, as the reference's
 *   batch_eval -> target_distribution chain does.  Labels always come from the
 *   unrounded q (first index wins on ties).
 * q            [n, K] float32 out, or NULL (fused latent-buffer mode writes nothing n-sized)
 * labels       [n] int32 out, or NULL
 * labels_prev  [n] int32 in, or NULL (then n_label_changes = 0)
 * stats        [K+1] float64 out: f[0..K), number of labels that differ from labels_prev
 */
int scc_dec_assign(const float* z, int64_t n, int d,
                   const float* mu, int K, float alpha, int round_decimals,
                   float* q, int32_t* labels, const int32_t* labels_prev,
                   double* stats, void* workspace, size_t workspace_bytes,
                   scc_stream_t stream);

/*
 * Target distribution  p_ij = (q_ij^2 / f_j) / sum_j (q_ij^2 / f_j)
 *   replaces Cluster/models.py:1320-1322 (target_distribution) given the
 *   column sums f from scc_dec_assign (after any cross-GPU allreduce).
 * round_decimals: 0, or 5 for the reference's np.round(p, 5).
 * A column with f_j == 0 yields NaN rows exactly like the reference.
 */
int scc_dec_target(const float* q, int64_t n, int K, const double* f,
                   int round_decimals, float* p, scc_stream_t stream);

/*
 * Column sums f_j = sum_i q_ij of an existing [n, K] matrix (models.py:1320, for
 * callers that hand target_distribution a q they did not get from scc_dec_assign).
 * workspace: at least scc_workspace_bytes(4, K) bytes.
 */
int scc_colsum(const float* q, int64_t n, int K, double* f,
               void* workspace, size_t workspace_bytes, scc_stream_t stream);

/*
 * KL(P||Q) loss and its gradients, q recomputed from z (never re-read):
 *   loss   = scale * sum_ij p_ij (log p_ij - log q_ij)      (terms with p_ij == 0 are 0)
 *   dz_i   =  scale (alpha+1)/alpha sum_j (p_ij - q_ij s_i) u_ij (z_i - mu_j)
 *   dmu_j  = -scale (alpha+1)/alpha sum_i (p_ij - q_ij s_i) u_ij (z_i - mu_j)
 *   replaces `gamma * KLDivLoss('sum')(log q, p) / B` + autograd backward,
 *   Cluster/models.py:1124-1127 (scale = gamma / B).
 *
 * p  [n, K] float32 target (API mode), or NULL: then p is rebuilt per point from
 *    f_cols[K] exactly as scc_dec_assign + scc_dec_target would have produced it
 *    (same round_decimals on q and on p) — the fused latent-buffer mode.
 * dz [n, d] float32 out, or NULL (centroid-only refinement)
 * stats [K*d + 2] float64 out: loss, sum_i s_i, dmu (row-major [K, d])
 */
int scc_dec_kl_grad(const float* z, int64_t n, int d,
                    const float* mu, int K, float alpha,
                    const float* p, const double* f_cols, int round_decimals,
                    float scale, float* dz, double* stats,
                    void* workspace, size_t workspace_bytes, scc_stream_t stream);

/*
 * Backward of the layer for an arbitrary upstream gradient G = dL/dq
 * (the autograd path of the literal reference loop, models.py:1122-1127):
 *   dz [n, d] float32 out or NULL; stats [K*d + 2] float64 out: 0, 0, dmu[K*d].
 */
int scc_dec_backward(const float* z, int64_t n, int d,
                     const float* mu, int K, float alpha,
                     const float* grad_q, float* dz, double* stats,
                     void* workspace, size_t workspace_bytes, scc_stream_t stream);

/*
 * One Lloyd (k-means) step: nearest-centre labels, per-cluster counts, centre
 * shifts and the inertia — the scan KMeans(n_init=100, max_iter=1000) repeats
 * inside Cluster/models.py:386-394 (seeding of gmm()) and models.py:565-574.
 *   stats [K*d + 2 + K] float64 out: inertia = sum_i min_j ||z_i - c_j||^2, 0,
 *         shift[K*d] = sum_{i in j} (z_i - c_j)  (new centre = c_j + shift_j / count_j),
 *         count[K]
 *   labels  [n] int32 out or NULL;  mindist [n] float32 out or NULL (k-means++ sampling)
 */
#define SCC_KMEANS_STATS(K, d) ((K) * (d) + 2 + (K))
int scc_kmeans_step(const float* z, int64_t n, int d,
                    const float* centers, int K,
                    int32_t* labels, float* mindist, double* stats,
                    void* workspace, size_t workspace_bytes, scc_stream_t stream);

/*
 * Batched Lloyd iterations — the `n_init` restarts of KMeans(n_clusters, n_init=100, max_iter=1000)
 * (Cluster/models.py:386-394, 565-572) advance TOGETHER: one launch scans z against restarts x K centres
 * (grid.y = restart), then scc_kmeans_batch_update moves every restart's centres and decides its
 * convergence on the device; the host polls the `done` flags every few iterations instead of once per
 * restart and iteration.
 *   centers [restarts, K, d] float32;  stats [restarts, K*d + 2 + K] float64 (layout of scc_kmeans_step)
 *   done    [restarts] uint8 or NULL: restarts whose flag is set are skipped (their stats are left untouched)
 *   labels  [restarts, n] int32 or NULL;  mindist [restarts, n] float32 or NULL (k-means++ sampling)
 *   workspace: scc_kmeans_batch_workspace_bytes(d, K, restarts) bytes, zeroed once (scc_workspace_init)
 * scc_kmeans_batch_update: c_j += shift_j / count_j (an empty cluster keeps its centre), n_iter += 1,
 *   inertia[r] = stats[r][0], done[r] = 1 once sum_j ||shift_j / count_j||^2 <= shift_tol
 *   (scikit-learn's stop rule: shift_tol = tol * mean feature variance).  n_iter, inertia nullable.
 */
size_t scc_kmeans_batch_workspace_bytes(int d, int K, int restarts);
int scc_kmeans_batch_step(const float* z, int64_t n, int d, const float* centers, int K, int restarts,
                          const unsigned char* done, int32_t* labels, float* mindist, double* stats,
                          void* workspace, size_t workspace_bytes, scc_stream_t stream);
int scc_kmeans_batch_update(float* centers, const double* stats, int d, int K, int restarts, double shift_tol,
                            unsigned char* done, int32_t* n_iter, double* inertia, scc_stream_t stream);

/*
 * Distance scan  D_ij = (sum_c |z_ic - mu_jc|^p)^(1/p),  out [n, K] float32 — the scan behind
 * utils.fractional_distance / utils.distance_matrix (Cluster/utils.py:866-869, 635-643) that the
 * reference's CDF/PDF/inertia analyses repeat per centroid (plotting.py:189,243,352; p = 2 gives the
 * Euclidean distances of utils.measure_class_inertia, utils.py:1024-1029).  Any d in [1,32], K in [1,16], p > 0.
 */
int scc_dec_distances(const float* z, int64_t n, int d, const float* mu, int K, float p, float* out,
                      scc_stream_t stream);

/* ------------------------------------------------------------------------- *
 * DEC stage, float64 precision path
 * ------------------------------------------------------------------------- *
 * The reference runs DEC in float64 (`model.double()`, Cluster/models.py:965; numpy float64 in batch_eval and
 * target_distribution, models.py:66-71, 1320-1322).  These entry points evaluate the same operators in IEEE float64
 * in the reference's operation order (any d in [1,32], K in [1,16]; no alignment requirement), so a caller that
 * keeps the reference's dtype gets its numbers to ~1e-15 and the np.round(., 5) quantisations land on the same
 * side.  Throughput is ~10x below the float32 kernels (one thread per point, FP64 pipe).
 *   scc_dec_assign_f64   as scc_dec_assign  (networks.py:279-288, models.py:92-94,1098-1099,1320)
 *   scc_dec_target_f64   as scc_colsum + scc_dec_target (models.py:1320-1322); have_f = 0: f [K] is computed from q
 *                        first (and returned), have_f = 1: f is given
 *   scc_dec_grad_f64     exactly one of p (target, API mode), f_cols (target rebuilt from the column sums, optionally
 *                        written to p_out) or grad_q (generic upstream gradient dL/dq) is non-NULL;
 *                        stats [K*d+2] = loss, sum_i s_i, dmu (models.py:1124-1127 + autograd)
 */
int scc_dec_assign_f64(const double* z, int64_t n, int d, const double* mu, int K, double alpha, int round_decimals,
                       double* q, int32_t* labels, const int32_t* labels_prev, double* stats,
                       void* workspace, size_t workspace_bytes, scc_stream_t stream);
int scc_dec_target_f64(const double* q, int64_t n, int K, double* f, int have_f, int round_decimals, double* p,
                       void* workspace, size_t workspace_bytes, scc_stream_t stream);
int scc_dec_grad_f64(const double* z, int64_t n, int d, const double* mu, int K, double alpha, const double* p,
                     const double* f_cols, int round_decimals, const double* grad_q, double scale, double* p_out,
                     double* dz, double* stats, void* workspace, size_t workspace_bytes, scc_stream_t stream);

struct scc_exchange;   /* multi-GPU exchange descriptor, defined below */

/* ------------------------------------------------------------------------- *
 * GMM stage (full covariance)
 * ------------------------------------------------------------------------- */

/* Packed float32 parameter block the E-step reads (written by scc_gmm_finalize
 * or scc_gmm_pack_params): means[K*d], U[K*d*(d+1)/2] (upper-triangular
 * precision-Cholesky factor, column-packed: U[a][b], a<=b at b(b+1)/2+a),
 * cst[K] = log det U_k + log pi_k - d/2 log(2 pi). */
#define SCC_GMM_TRI(d) ((d) * ((d) + 1) / 2)
#define SCC_GMM_PARAM_FLOATS(K, d) ((K) * (d) + (K) * SCC_GMM_TRI(d) + (K))
/* Packed float64 statistics of one fused E+M pass:
 * [0] sum_i log p(x_i); then per k: N_k ; then S1[K*d] = sum_i r_ik (x_i - mu_k);
 * then S2[K*TRI] = sum_i r_ik (x_i-mu_k)_a (x_i-mu_k)_b, a<=b column-packed. */
#define SCC_GMM_STAT_DOUBLES(K, d) (1 + (K) + (K) * (d) + (K) * SCC_GMM_TRI(d))
/* Control block (float64[8]) shared by finalize calls of one fit:
 * [0] lower bound of the last E-step, [1] previous lower bound, [2] n_iter,
 * [3] converged flag, [4] not-positive-definite flag (index k+1 of first bad
 * component), [5] frozen flag (set with converged: later passes are no-ops). */
#define SCC_GMM_CTRL_DOUBLES 8

/*
 * One fused E-step + M-step sufficient-statistics pass over z:
 *   replaces sklearn:mixture/_base.py:314-332,552-582 (_e_step, log-sum-exp
 *   responsibilities), sklearn:mixture/_gaussian_mixture.py:490-553 (Cholesky
 *   log-likelihoods) and the sums of :282-320,168-197 (_m_step) — the calls
 *   Cluster/models.py:411 (GMM.fit_predict) spends its time in.
 * Moments are centred on the CURRENT means (the ones in `params`), so the
 * finalize step applies mu_new = mu + S1/N_k, Sigma = S2/N_k - dd^T + reg I.
 * labels [n] int32 out or NULL: argmax_k of the responsibilities (final pass).
 * resp   [n, K] float32 out or NULL: the responsibilities themselves.
 * ctrl   float64[8] or NULL: if ctrl[5] != 0 the pass is skipped (frozen fit).
 * mode: SCC_GMM_ESTEP_ONLY skips the M-step sums (label-only final E-step);
 *   SCC_GMM_SOFT is the EM pass; SCC_GMM_HARD replaces r_ik by the one-hot
 *   argmax_k (the initial responsibilities sklearn builds from k-means labels,
 *   sklearn:mixture/_base.py:119-128, when params hold Sigma = I, pi = 1/K).
 */
#define SCC_GMM_ESTEP_ONLY 0
#define SCC_GMM_SOFT 1
#define SCC_GMM_HARD 2
/* OR-able flag: keep every (point, component) pair in the M-step sums.  By default pairs whose responsibility
 * is below 2^-30 are skipped (they move no moment by more than 1e-9 relative; after the first EM iterations
 * that is ~90 % of all pairs at K = 16) — the d <= 12 kernels build compact per-component point lists. */
#define SCC_GMM_NOSKIP 16
int scc_gmm_em_step(const float* z, int64_t n, int d, int K,
                    const float* params, double* stats,
                    int32_t* labels, float* resp, const double* ctrl,
                    int mode,
                    void* workspace, size_t workspace_bytes, scc_stream_t stream);

/*
 * M-step finalisation on the device (one thread block; float64):
 *   N_k += 10 eps; pi = N_k / sum N_k; mu, Sigma (+reg_covar on the diagonal);
 *   L = chol(Sigma); U = L^-T; log det; lower bound = stats[0] / n_total;
 *   converged if |lower - previous| < tol.
 *   replaces sklearn:mixture/_gaussian_mixture.py:883-901, 282-320 (divisions),
 *   323-385 (_compute_precision_cholesky), 448-487 and the convergence test of
 *   sklearn:mixture/_base.py:270-278.
 * stats        float64 packed sums (after any cross-GPU allreduce)
 * means        float64 [K, d] in/out (current means in, new means out)
 * weights      float64 [K] out;  covariances float64 [K, d, d] out;
 * prec_chol    float64 [K, d, d] out (upper triangular, sklearn's precisions_cholesky_)
 * params       float32 packed block for the next E-step, out
 * ctrl         float64[8] in/out (see SCC_GMM_CTRL_DOUBLES)
 */
int scc_gmm_finalize(const double* stats, double n_total, int d, int K,
                     double reg_covar, double nk_eps, double tol,
                     double* means, double* weights, double* covariances,
                     double* prec_chol, float* params, double* ctrl,
                     scc_stream_t stream);

/*
 * One whole EM iteration in two launches: the fused E+M statistics pass over z (scc_gmm_em_step) and ONE tail
 * kernel that sums the per-thread-block partial statistics in a fixed order, all-reduces them over the GPUs
 * (`exchange`, or NULL on a single GPU: every thread block ships / polls its own slice of the vector through the
 * flag-in-data exchange window) and runs the M-step finalisation of scc_gmm_finalize in its last thread block.
 * Same arithmetic as scc_gmm_em_step -> (all-reduce) -> scc_gmm_finalize; `stats` receives the world's sums.
 * mode: SCC_GMM_SOFT or SCC_GMM_HARD, optionally | SCC_GMM_NOSKIP.  A frozen fit (ctrl[5] != 0) is a no-op.
 * Replaces one trip of the loop in sklearn:mixture/_base.py:262-278.
 */
int scc_gmm_em_iteration(const float* z, int64_t n, int d, int K, float* params, double* stats, int mode,
                         double n_total, double reg_covar, double nk_eps, double tol,
                         double* means, double* weights, double* covariances, double* prec_chol, double* ctrl,
                         void* workspace, size_t workspace_bytes, const struct scc_exchange* exchange,
                         scc_stream_t stream);

/*
 * Build the packed E-step parameter block from explicit (pi, mu, Sigma) float64
 * device arrays — the initial state (weights_init / means_init + initial
 * covariances, sklearn:mixture/_gaussian_mixture.py:848-881).  Also fills
 * prec_chol [K,d,d] (may be NULL) and resets ctrl.
 */
int scc_gmm_pack_params(const double* weights, const double* means,
                        const double* covariances, int d, int K,
                        double* prec_chol, float* params, double* ctrl,
                        scc_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Multi-GPU exchange of the packed statistics (one process per GPU, one NVSwitch box)
 * ------------------------------------------------------------------------- */

/*
 * One-shot all-reduce(sum) of a float64 vector over NVLink peer memory, in one small kernel
 * (the reference has no multi-GPU path; this is the exchange step SURVEY.md 8e names: f_j,
 * loss + dL/dmu, N_k / sum r z / sum r zz^T).
 *   peer_windows: DEVICE array of `world` pointers, entry r = rank r's exchange window mapped into
 *                 this process (symmetric allocation of scc_peer_window_bytes(max_len) bytes,
 *                 zero-filled once before first use, same max_len on every rank).
 *   Every rank must call with the same len in the same order; the result (summed in rank order)
 *   is bit-identical on all ranks.  local and out may alias.
 */
size_t scc_peer_window_bytes(int max_len);

/*
 * Fused form: the exchange rides on the kernels that produce / consume the statistics, so a sharded
 * DEC step is  assign_ex(push) -> target_kl_grad(pull_f, push)  — two launches, no separate collective.
 * `push` on scc_dec_assign_ex: the LAST thread block of the kernel ships the reduced vector to every rank's
 * window (peer stores over NVLink in the same kernel as the compute; vectors of <= 1024 doubles travel
 * flag-in-data, longer ones behind a system fence + sequence flag); it must be matched by exactly one pull
 * on every rank (target_ex / kl_grad_ex / target_kl_grad with pull_f, or scc_peer_finish) before the
 * next-but-one push.  `pull`: the consuming kernel waits for the world's vectors in its prologue and sums
 * them in rank order.  `push` on the GRADIENT kernels (scc_dec_kl_grad_ex, scc_dec_target_kl_grad,
 * scc_dec_step_ex) is a complete all-reduce: the last thread block pushes, waits for the world and overwrites
 * `stats` with the rank-ordered sum before the kernel ends.
 */
typedef struct scc_exchange {
    void* const* windows;   /* DEVICE array of `world` window pointers (see scc_peer_allreduce) */
    int rank, world, max_len;
} scc_exchange;

int scc_dec_assign_ex(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                      float* q, int32_t* labels, const int32_t* labels_prev, double* stats,
                      void* workspace, size_t workspace_bytes, const scc_exchange* push, scc_stream_t stream);
/* f [K+1] receives the all-reduced column sums and label-change count when `pull` is given */
int scc_dec_target_ex(const float* q, int64_t n, int K, double* f, int round_decimals, float* p,
                      const scc_exchange* pull, scc_stream_t stream);
int scc_dec_kl_grad_ex(const float* z, int64_t n, int d, const float* mu, int K, float alpha,
                       const float* p, const double* f_cols, int round_decimals, float scale, float* dz,
                       double* stats, void* workspace, size_t workspace_bytes,
                       const scc_exchange* pull_f, const scc_exchange* push, scc_stream_t stream);
/*
 * scc_dec_target_kl_grad — target distribution + KL loss + gradients in ONE pass over z.
 *   The step `p = target_distribution(np.round(q, 5))` (Cluster/models.py:1302-1322) followed by
 *   `gamma * KLDivLoss('sum')(log q, p) / B` + backward (models.py:1124-1127) for the same batch:
 *   q is recomputed from z in registers, p is rebuilt per point from the column sums f_cols[K]
 *   (output of scc_dec_assign; or pulled from the exchange `pull_f`) with the same rounding as
 *   scc_dec_target, consumed by the loss / gradient and — when p_out != NULL — written out as the
 *   [n, K] float32 target the reference keeps (`tar_dist`).  Replaces scc_dec_target +
 *   scc_dec_kl_grad(p) and saves one read of q, one read of p and a launch.
 *   p_out, dz nullable; pull_f / push nullable (single GPU).  stats as scc_dec_kl_grad.
 */
int scc_dec_target_kl_grad(const float* z, int64_t n, int d, const float* mu, int K, float alpha,
                           const double* f_cols, int round_decimals, float scale, float* p_out, float* dz,
                           double* stats, void* workspace, size_t workspace_bytes,
                           const scc_exchange* pull_f, const scc_exchange* push, scc_stream_t stream);
/*
 * Two-launch DEC step for the shapes the one-kernel step does not cover (K*d > 160, e.g. the d = 32, K = 16 shard of
 * BASELINE configs[3]), with a hand-off between the passes: scc_dec_assign_u also writes u_ij = 1 / (1 + d_ij / alpha)
 * ([n, K] float32, the Student's-t kernel before normalisation) and scc_dec_target_kl_grad_u streams it back instead of
 * recomputing the K*d distance terms — at that shape both passes are bound by the shared-memory reads of the centroid
 * table (2 KB per point), not by HBM, so 2 x 4K extra bytes per point through HBM are the cheaper way.
 * Other arguments as scc_dec_assign_ex / scc_dec_target_kl_grad.  Returns SCC_ERR_UNSUPPORTED for the shapes the
 * one-kernel step serves (use scc_dec_step there).
 */
int scc_dec_assign_u(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                     float* q, int32_t* labels, const int32_t* labels_prev, float* u_out, double* stats,
                     void* workspace, size_t workspace_bytes, const scc_exchange* push, scc_stream_t stream);
int scc_dec_target_kl_grad_u(const float* z, int64_t n, int d, const float* mu, int K, float alpha,
                             const float* u_in, const double* f_cols, int round_decimals, float scale,
                             float* p_out, float* dz, double* stats, void* workspace, size_t workspace_bytes,
                             const scc_exchange* pull_f, const scc_exchange* push, scc_stream_t stream);

/*
 * scc_dec_step — the whole DEC step of one batch in ONE kernel (single GPU):
 *   pass 1 = scc_dec_assign (q, labels, f, label-change count), a grid-wide barrier that all-reduces f
 *   across the CTAs, pass 2 = scc_dec_target_kl_grad (p, loss, dL/dz, dL/dmu) with z re-read from L2.
 *   Same arithmetic as the two stand-alone kernels (q and p bit-identical to them).  Cooperative launch:
 *   all CTAs are co-resident.  Replaces networks.py:279-288 + models.py:92-94,1098-1099,1302-1322,1124-1127
 *   for a batch that is processed whole (batch_eval + the update step on the same latent set).
 *   q, labels, labels_prev, p_out, dz nullable.  f_stats [K+1], stats [K*d+2] float64 out.
 *   Returns SCC_ERR_UNSUPPORTED for (d, K) served by the tiled gradient kernel (K*d > 160): call the two
 *   kernels instead.
 */
int scc_dec_step(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                 float scale, float* q, int32_t* labels, const int32_t* labels_prev, double* f_stats,
                 float* p_out, float* dz, double* stats, void* workspace, size_t workspace_bytes,
                 scc_stream_t stream);
/* Multi-GPU form: BOTH all-reduces run INSIDE the kernel — the last CTA to reach the grid barrier between the
 * two passes pushes this GPU's f to every rank's exchange window (NVLink peer stores, flag-in-data), waits for
 * the world's vectors and releases the other CTAs with the rank-ordered sum; the kernel's last CTA does the
 * same with the final statistics, so `stats` holds the world's loss / dL/dmu when the kernel ends (one launch
 * per step and GPU, no separate collective).  scale = gamma / N_total; f_stats receives the all-reduced f.
 * Every rank must launch it (an empty shard, n == 0, included). */
int scc_dec_step_ex(const float* z, int64_t n, int d, const float* mu, int K, float alpha, int round_decimals,
                    float scale, float* q, int32_t* labels, const int32_t* labels_prev, double* f_stats,
                    float* p_out, float* dz, double* stats, void* workspace, size_t workspace_bytes,
                    const scc_exchange* ex, scc_stream_t stream);
int scc_peer_finish(double* out, int len, const scc_exchange* ex, scc_stream_t stream);
int scc_peer_allreduce(const double* local, int len, double* out,
                       void* const* peer_windows, int rank, int world, int max_len,
                       scc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SCC_B200_H_ */
