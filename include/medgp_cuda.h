/*
 * medgp_cuda.h -- C ABI of libmedgp_cuda.so, the B200 (sm_100a) backend for MedGP's
 * per-patient exact-inference hot path.
 *
 * The reference (bee-hive/MedGP) has no FFI; its seam is C++ virtual dispatch.  Each entry
 * point below names the reference interface it replaces (paths relative to the reference
 * root, medgpc/src/...).  All pointers are HOST pointers unless the name starts with d_;
 * the library owns all device memory; no exception crosses the boundary; every function
 * returns MEDGP_OK (0) or a negative medgp_status.  One context drives ONE GPU (one process
 * or host thread per GPU; contexts are independent, so a cohort shards over GPUs with no
 * collective).  A context is thread-compatible, not thread-safe.
 *
 * Hyper-parameter vector theta (length P = D + Q*(D*R+2+D), doubles), exactly the
 * reference's flat order (core/c_hyperparam.cpp:99-121, kernel/c_kernel_LMC_SM.cpp:51-62):
 *   [ log sigma_d (D) | A_q[d][r] raw (Q*D*R, index q*D*R+d*R+r) | log mu_q (Q) |
 *     log v_q (Q) | log kappa_q[d] (Q*D, index q*D+d) ]
 * Gradients come back in the same order, w.r.t. the stored (log or raw) values, WITHOUT
 * prior terms (those stay on the host: inference/c_inference_prior.cpp:59-150).
 */
#ifndef MEDGP_CUDA_H
#define MEDGP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct medgp_ctx medgp_ctx;

typedef enum {
    MEDGP_OK = 0,
    MEDGP_ERR_ARG = -1,      /* bad argument (NULL, out of range, model not set) */
    MEDGP_ERR_CUDA = -2,     /* a CUDA call failed; see medgp_cuda_last_error   */
    MEDGP_ERR_NOMEM = -3,    /* workspace too small for even one evaluation      */
    MEDGP_ERR_NODEVICE = -4  /* no usable sm_100 device                          */
} medgp_status;

/* Per-evaluation status written to the `status` arrays:
 *   0   success, no jitter;  k>0  success after k extra noise additions to the diagonal
 *   -1  still not positive definite after 10 additions (the reference returns false:
 *       inference/c_inference_exact.cpp:99-111) -- nlml/grad/mean/var are then NaN. */

/* Stages for which device time is accumulated (CUDA events on the library's stream). */
enum {
    MEDGP_STAGE_PREP = 0,     /* theta -> B_q, sigma^2, per-point cos/sin tables          */
    MEDGP_STAGE_ASSEMBLE,     /* kernel (1): covariance assembly                           */
    MEDGP_STAGE_POTRF,        /* kernel (2): blocked FP64 Cholesky, panel/update kernels (DMMA) */
    MEDGP_STAGE_DIAG,         /* kernel (2): 64x64 diagonal-block Cholesky + inverse       */
    MEDGP_STAGE_SOLVE,        /* kernel (2): triangular solves, log-det, NLML              */
    MEDGP_STAGE_TRTRI,        /* kernel (2): L^-1                                          */
    MEDGP_STAGE_LAUUM,        /* kernel (2): K^-1 = L^-T L^-1                              */
    MEDGP_STAGE_GRAD,         /* kernel (3): fused gradient reduction                      */
    MEDGP_STAGE_PREDICT,      /* kernel (4): cross-covariance, predictive mean/variance    */
    MEDGP_STAGE_COUNT
};

typedef struct {
    double ms[MEDGP_STAGE_COUNT];        /* accumulated device milliseconds per stage        */
    long long launches[MEDGP_STAGE_COUNT]; /* kernel launches per stage                      */
    double flops[MEDGP_STAGE_COUNT];     /* algorithmic FP64 flop     */
    double bytes[MEDGP_STAGE_COUNT];     /* algorithmic HBM bytes                            */
    long long evals;                     /* evaluations completed                            */
} medgp_stage_times;

/* Create a context on CUDA device `device`.  workspace_bytes = 0 picks 70% of free HBM.
 * Fails with MEDGP_ERR_NODEVICE when there is no GPU: there is NO CPU fallback. */
int medgp_cuda_create(medgp_ctx **out, int device, size_t workspace_bytes);
void medgp_cuda_destroy(medgp_ctx *ctx);
const char *medgp_cuda_last_error(const medgp_ctx *ctx);

/* Kernel shape.  Replaces c_kernel_LMC_SM::set_kernel_param (kernel/c_kernel_LMC_SM.cpp:64-70)
 * and c_likelihood_gaussianMO's output count (likelihoods/c_likelihood_gaussianMO.cpp:25-29).
 * pi_const is the reference's truncated PI, 3.14159265 (util/global_settings.h:6). */
int medgp_cuda_model(medgp_ctx *ctx, int Q, int D, int R, double pi_const);
int medgp_cuda_num_hyp(const medgp_ctx *ctx);  /* P, or negative status */

/* Upload one series (a patient, or one sliding-window training set) once.  Replaces the
 * data copies c_objective_one keeps (util/c_objective_one.cpp:23-36).  meta[i] in [0,D) is
 * the feature slot of point i (dataio/c_experiment.cpp:298), x hours, y z-scored value.
 * Any point order is accepted; results do not depend on it. */
int medgp_cuda_add_series(medgp_ctx *ctx, int n, const int32_t *meta, const float *x,
                          const float *y, int *out_series_id);
/* Internal point order of an uploaded series.  FEATURE (what medgp_cuda_add_series uses) serves
 * NLML, gradients and medgp_cuda_predict; TIME serves medgp_cuda_predict_online (and NLML /
 * medgp_cuda_predict), not gradients.  At most 32 points may share one timestamp with TIME. */
enum { MEDGP_ORDER_FEATURE = 0, MEDGP_ORDER_TIME = 1, MEDGP_ORDER_GIVEN = 2 };
/* GIVEN keeps the caller's point order (what medgp_cuda_export_factors refers to); NLML and
 * medgp_cuda_predict work on it, gradients only if that order is feature-major already. */
int medgp_cuda_add_series_ordered(medgp_ctx *ctx, int n, const int32_t *meta, const float *x,
                                  const float *y, int order, int *out_series_id);
/* `count` series in one call -- one device allocation, one host-to-device copy: n[b] points each,
 * meta / x / y concatenated in batch order.  For callers that upload a training window per patient
 * and time stamp (the with-update branch of run_test_one, main_one_test.cpp:286-348). */
int medgp_cuda_add_series_batch(medgp_ctx *ctx, int count, const int *n, const int32_t *meta, const float *x,
                                const float *y, int order, int *out_series_ids);
int medgp_cuda_free_series(medgp_ctx *ctx, int series_id);
int medgp_cuda_free_series_batch(medgp_ctx *ctx, int count, const int *series_ids);
int medgp_cuda_clear_series(medgp_ctx *ctx);

/* batch NLML (+ gradient) evaluations: evaluation b uses series_id[b] and theta[b*P..].
 * Replaces c_objective_one::compute_objective -> GP_Regression::train ->
 * c_inference_exact::compute_nlml (util/c_objective_one.cpp:40-81, core/gp_regression.cpp:102-126,
 * inference/c_inference_exact.cpp:29-244) including kernel/c_kernel_LMC_SM.cpp:152-327.
 * grad may be NULL when want_grad == 0. */
int medgp_cuda_nlml_grad(medgp_ctx *ctx, int batch, const int *series_id, const double *theta,
                         int want_grad, double *nlml, double *grad, int *status);

/* Same work with theta already resident in HBM and results left there (d_theta: batch*P,
 * d_nlml: batch, d_grad: batch*P or NULL, d_status: batch ints); asynchronous on the
 * context's stream, with no host synchronisation inside the call -- call medgp_cuda_sync
 * before reading.  Jitter retries (inference/c_inference_exact.cpp:99-108) run on the device:
 * the launch sequence of a chunk is the body of a CUDA-graph WHILE node that re-runs the failed
 * evaluations with one more noise addition until none is left, so status is final (0, k, or -1
 * after 10 additions) exactly as on the host path. */
int medgp_cuda_nlml_grad_device(medgp_ctx *ctx, int batch, const int *series_id,
                                const double *d_theta, int want_grad, double *d_nlml,
                                double *d_grad, int *d_status);
int medgp_cuda_sync(medgp_ctx *ctx);

/* batch predictions: evaluation b trains on series_id[b] with theta[b*P..] and predicts the
 * n_star[b] points meta_star/x_star[star_offset[b] ..].  Replaces GP_Regression::train(false)
 * + GP_Regression::predict (core/gp_regression.cpp:102-214; cross-covariance
 * kernel/c_kernel_LMC_SM.cpp:329-372, prior variance :122-150).  mean/var are indexed like
 * x_star.  star_offset has batch+1 entries. */
int medgp_cuda_predict(medgp_ctx *ctx, int batch, const int *series_id, const double *theta,
                       const int *star_offset, const int32_t *meta_star, const float *x_star,
                       double *mean, double *var, int *status);

/* Online one-step-ahead imputation of whole series, the loop of run_test_one without
 * hyper-parameter updates (main_one_test.cpp:269-444), with ONE factorisation per series instead
 * of one per observation: for every point j of series_id[b] (uploaded with MEDGP_ORDER_TIME),
 * mean/var of y_j given all points with an earlier timestamp plus the other points sharing j's
 * timestamp (:286-306, :354-366), under theta[b*P..].  var includes the noise of j's feature
 * (core/gp_regression.cpp:185-196).  Outputs are concatenated in batch order, each series'
 * n values in the caller's point order.  A point with no training data gets mean 0 and its prior
 * variance.  No jitter on this path: status[b] = -1 (results NaN) when the time-ordered
 * matrix is not positive definite -- the caller then falls back to medgp_cuda_predict per
 * observation, which retries with jitter exactly as the reference does. */
int medgp_cuda_predict_online(medgp_ctx *ctx, int batch, const int *series_id, const double *theta,
                              double *mean, double *var, int *status);

/* The out-parameters of c_inference::compute_nlml for callers that keep the reference's own
 * GP_Regression::predict (core/gp_regression.cpp:169-196): chol_alpha = K^-1 y (n floats) and
 * chol_factor_inv = L^-1 (n*n floats, row-major, lower, strict upper part zero:
 * inference/c_inference_exact.cpp:124-143), plus the NLML and the jitter status, for a series
 * uploaded with MEDGP_ORDER_GIVEN.  One factorisation + triangular inverse on the GPU, then an
 * n^2 read-back; the batched medgp_cuda_predict / medgp_cuda_predict_online are the fast path. */
int medgp_cuda_export_factors(medgp_ctx *ctx, int series_id, const double *theta, float *alpha,
                              float *Linv, double *nlml, int *status);

/* Debug/parity taps (tests only; one evaluation): the assembled K+noise (n*n, row-major,
 * full symmetric) for kernel (1); the Cholesky factor L (n*n row-major lower) and
 * alpha = K^-1 y for kernel (2); K^-1 (n*n row-major full) after want_grad.
 * All in the caller's original point order. */
int medgp_cuda_debug_matrices(medgp_ctx *ctx, int series_id, const double *theta,
                              double *K, double *L, double *alpha, double *Kinv);

/* ---- Device-resident lock-step optimiser.  Replaces the loop of c_optimizer_scg::optimize
 * (util/c_optimizer_scg.cpp:25-284: Rasmussen's minimize, Polak-Ribiere CG with cubic /
 * quadratic line search) for `count` instances at once -- one per (patient, initialisation) --
 * together with the prior terms of c_inference_prior::compute_nlml
 * (inference/c_inference_prior.cpp:59-150).  The line-search state of every instance lives in
 * HBM; one super-step = one batched NLML+gradient evaluation at every live instance's probe
 * point + one kernel that advances all state machines; theta and gradients never cross PCIe,
 * the host only polls how many instances still want evaluations.  Control flow is the
 * reference's, evaluation for evaluation (summation order of the dot products differs). */
typedef struct medgp_scg medgp_scg;
int medgp_cuda_scg_create(medgp_ctx *ctx, int count, medgp_scg **out);
void medgp_cuda_scg_destroy(medgp_scg *scg);
/* (Re)start every instance: instance b optimises series_id[b] from theta0[b*P..] with budget
 * max_iteration[b] < 0 = -(function evaluations), the only form MedGP uses (main_one_train.cpp:270-291,
 * c_optimizer_varEM.cpp:64-69); a budget >= 0 leaves the instance finished at theta0.
 * Priors, per instance and hyper-parameter in theta order (all three arrays or none):
 *   prior_type  count*P: -1 none, 0 clamp (gradient forced to 0), 1 normal(p0, VARIANCE p1),
 *               2 laplace(p0, scale p1)                          (prior/c_prior.cpp:383-421)
 *   prior_exp   count*P: 1 = the hyper-parameter is stored as a log, d log p gets the factor h
 *   prior_param count*P*2: (p0, p1) */
int medgp_cuda_scg_start(medgp_scg *scg, const int *series_id, const double *theta0, const int *max_iteration,
                         const signed char *prior_type, const signed char *prior_exp, const float *prior_param);
/* Enqueue `super_steps` super-steps (finished instances are passed over on the device), wait,
 * and report how many instances still want evaluations. */
int medgp_cuda_scg_run(medgp_scg *scg, int super_steps, int *active_left);
/* Best point, its objective (NLML + prior terms) and the evaluations spent, per instance
 * (opt_parameter / opt_loss of c_optimizer_scg::optimize).  Any pointer may be NULL. */
int medgp_cuda_scg_result(medgp_scg *scg, double *theta_best, double *loss, int *evals);
/* Test taps: drive the SAME device state machine with an external objective -- fetch every
 * instance's probe point (wants[b] = 0 once finished), then feed (f, gradient, ok) for all. */
int medgp_cuda_scg_points(medgp_scg *scg, double *theta, int *wants);
int medgp_cuda_scg_feed(medgp_scg *scg, const double *f, const double *grad, const int *ok);

/* ---- Population "mode kernel" (between training and testing): batched Gaussian-KDE mode
 * estimation.  Replaces compute_kde + compute_mode(weighted=True) of
 * medgpc/clustering/mode_estimate.py:438-450 (statsmodels KDEUnivariate, kernel "gau", evaluated at
 * the data points) for n_sets independent sets at once: set s holds data[offsets[s] .. offsets[s+1])
 * (offsets[0] = 0) with bandwidth[s] > 0 (the caller's Silverman rule); mode[s] = density-weighted
 * mean of the set; density (optional, like data) receives the per-point densities. */
int medgp_cuda_kde_mode(medgp_ctx *ctx, int n_sets, const int *offsets, const double *data,
                        const double *bandwidth, double *mode, double *density);

/* Tests of the jitter path: declare the first `attempts` factorisation attempts of every
 * evaluation failed, whatever their pivots (0 = off).  An evaluation then comes back with
 * status == attempts and the values of K + (1 + attempts) sigma^2 -- what the reference computes
 * when spotrf fails that many times.  MEDGP_FORCE_FAIL=<attempts> sets it at context creation
 * (for the executables). */
int medgp_cuda_debug_force_fail(medgp_ctx *ctx, int attempts);

/* Per-stage device timings and algorithmic work counters since the last reset. */
int medgp_cuda_profile(medgp_ctx *ctx, int enable);  /* enabling adds event records */
int medgp_cuda_stage_times(medgp_ctx *ctx, medgp_stage_times *out, int reset);

/* Raw device helpers so callers (bench, Python) can stage theta/results without torch. */
int medgp_cuda_malloc(medgp_ctx *ctx, size_t bytes, void **d_ptr);
int medgp_cuda_free(medgp_ctx *ctx, void *d_ptr);
int medgp_cuda_memcpy_h2d(medgp_ctx *ctx, void *d_dst, const void *src, size_t bytes);
int medgp_cuda_memcpy_d2h(medgp_ctx *ctx, void *dst, const void *d_src, size_t bytes);
/* Page-locked host memory.  medgp_cuda_nlml_grad copies straight from/to caller buffers that
 * are page-locked (allocated here, by cudaHostAlloc or registered with cudaHostRegister) and
 * stages pageable ones through its own pinned buffers. */
int medgp_cuda_host_alloc(medgp_ctx *ctx, size_t bytes, void **h_ptr);
int medgp_cuda_host_free(medgp_ctx *ctx, void *h_ptr);
/* The context's stream as a cudaStream_t (for event timing by the caller). */
void *medgp_cuda_stream(medgp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MEDGP_CUDA_H */
