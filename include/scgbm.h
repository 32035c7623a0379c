/*
 * scgbm.h -- C ABI of libscgbm, the B200-native replacement for the hot path of
 * phillipnicol/scGBM:  gbm.sc()  (R/fit.R:52-211, pgd_irlba :315-328,
 * process.results :301-313, gbm.proj.parallel :213-285)  and  get.se()
 * (R/uncertainty.R:14-35).
 *
 * The reference has no native boundary (pure R).  These are the entry points an
 * R `.Call` shim binds (r-package/src/shim.c) -- plain pointers and sizes, no R,
 * torch or CUDA types.  Conventions:
 *   - all matrices are COLUMN-MAJOR (R layout); genes = rows (I), cells = cols (J);
 *   - the caller owns every buffer it passes; the library owns device memory and
 *     frees it before returning, also on error;
 *   - every function returns 0 on success, non-zero on error; the message is
 *     retrievable with scgbm_last_error() (thread-local);
 *   - not re-entrant per device; the progress callback fires on the calling thread.
 *   - there is NO CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef SCGBM_H
#define SCGBM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCGBM_VERSION 200

/* element type of a count matrix handed to the library */
enum scgbm_dtype {
  SCGBM_I32 = 0, /* R integer matrix                    */
  SCGBM_F64 = 1, /* R double matrix                     */
  SCGBM_U8  = 2,
  SCGBM_U16 = 3,
  SCGBM_F32 = 4
};

/* device storage format of Y (bytes per stored count = s_Y in the roofline) */
enum scgbm_storage {
  SCGBM_STORE_AUTO = 0, /* u16 if max(Y) <= 65535 and integral, else f32 */
  SCGBM_STORE_U8   = 1,
  SCGBM_STORE_U16  = 2,
  SCGBM_STORE_F32  = 3
};

enum scgbm_status {
  SCGBM_OK = 0,
  SCGBM_ERR_INVALID = 1,   /* bad argument; mirrors stop() in R/fit.R:287-299,87 */
  SCGBM_ERR_CUDA = 2,      /* CUDA / NCCL runtime failure, or no device          */
  SCGBM_ERR_NUMERIC = 3,   /* e.g. NaN objective followed by NA comparison (Q1), */
                           /* singular information matrix in get.se              */
  SCGBM_ERR_NOMEM = 4
};

/* kernel classes for the built-in CUDA-event profiler (scgbm_profile.ms[k]) */
enum scgbm_kernel_class {
  SCGBM_K_PACK = 0,       /* K1 pack + integer stats (R/fit.R:94-96,102-105)      */
  SCGBM_K_ROWSUM = 1,     /* K3 alpha row sums        (R/fit.R:127-130)           */
  SCGBM_K_COLSUM = 2,     /* K3 beta col sums         (R/fit.R:133-135)           */
  SCGBM_K_FUSED = 3,      /* K4 fused likelihood pass (R/fit.R:136-151,321)       */
  SCGBM_K_PROD_N = 4,     /* K5 R.P   products        (R/fit.R:323)               */
  SCGBM_K_PROD_T = 5,     /* K5 R^T.Q products        (R/fit.R:323)               */
  SCGBM_K_SMALL = 6,      /* K6 small dense fp64 (Gram, eig, orth, rotate)        */
  SCGBM_K_SE = 7,         /* K8 get.se                                            */
  SCGBM_K_PROJ = 8,       /* K9 projection IRLS                                   */
  SCGBM_K_COMM = 9,       /* NCCL collectives                                     */
  SCGBM_K_MISC = 10,
  SCGBM_K_NCLASS = 11
};

typedef struct scgbm_profile {
  double  ms[SCGBM_K_NCLASS];        /* summed CUDA-event time per class            */
  int64_t launches[SCGBM_K_NCLASS];  /* kernel launches per class                   */
  double  fused_bytes;               /* algorithmic bytes of ONE fused-pass launch  */
  double  total_ms;                  /* whole call, device-timed                    */
  double  h2d_bytes, d2h_bytes;      /* host<->device traffic of the call           */
  int64_t svd_passes;                /* passes over the materialised residual       */
  int64_t svd_calls;
} scgbm_profile;

typedef struct scgbm_comm scgbm_comm; /* opaque: one NCCL rank */

typedef void (*scgbm_progress_cb)(int iter, double objective, void* user);

/* ---- gbm.sc(Y, M, ..., subset = NULL)   R/fit.R:52-211 ------------------------ */
typedef struct scgbm_fit_args {
  const void* Y;        /* I x J_local counts, column-major, leading dim ldY        */
  int32_t  y_dtype;     /* scgbm_dtype                                              */
  int32_t  y_on_device; /* 0: host pointer; 1: device pointer (stays resident)      */
  int64_t  I;           /* genes                                                    */
  int64_t  J;           /* cells held by THIS rank (all cells when comm == NULL)    */
  int64_t  ldY;         /* >= I                                                     */
  int32_t  M;           /* latent factors (R/fit.R:53)                              */
  int32_t  max_iter;    /* R default 100                                            */
  double   tol;         /* R default 1e-4                                           */
  int32_t  min_iter;    /* R default 30                                             */
  int32_t  infer_beta;  /* R default FALSE                                          */
  const int32_t* batch; /* J 1-based batch codes (as.integer(factor), R/fit.R:90) or NULL */
  int32_t  nbatch;      /* max(batch) over ALL ranks; 1 when batch == NULL          */
  int32_t  return_W;    /* fill out->W (I x J doubles)                              */
  int32_t  time_by_iter;/* fill out->time                                           */
  /* --- extensions (zero = default) --- */
  int32_t  y_storage;   /* scgbm_storage                                            */
  int32_t  device;      /* CUDA ordinal; <0: current device                         */
  scgbm_comm* comm;     /* NULL: single GPU; else cells are sharded over the ranks  */
  int32_t  svd_extra;   /* Krylov block = M + svd_extra   (default 12)              */
  int32_t  svd_depth;   /* max Krylov extensions per cycle (default 3)              */
  double   svd_tol;     /* Ritz residual estimate to stop at (0: 1e-5, irlba's default) */
  uint64_t seed;        /* start block of the initial SVD                           */
  int32_t  profile;     /* 1: per-class CUDA-event timing into out->prof; 2: only the */
                        /* fused pass (K4) is bracketed by events (launch counts stay) */
  int32_t  verbose;
  scgbm_progress_cb progress; /* called once per accepted iteration (R/fit.R:175)   */
  void*    progress_user;
  /* --- sparse input: used when Y == NULL.  Column-compressed like Matrix::dgCMatrix
   * (the reference densifies such input itself, R/fit.R:224); the dense device matrix is built by a
   * scatter kernel, so no I x J host copy ever exists. --- */
  const int32_t* csc_i; /* nnz 0-based row indices          (dgCMatrix@i)                       */
  const int64_t* csc_p; /* J + 1 column pointers            (dgCMatrix@p widened to 64 bit)     */
  const void*    csc_x; /* nnz values                       (dgCMatrix@x: SCGBM_F64)            */
  int32_t  csc_x_dtype; /* scgbm_dtype of csc_x                                                 */
  /* --- multi-GPU from ONE caller (the reference's only calling convention is a single R session calling
   * gbm.sc(), R/fit.R:52): ngpu >= 2 splits the cells over ngpu GPUs inside the library -- one host thread and
   * one NCCL rank per GPU, rank 0 on the calling thread (so `progress` still fires there).  Y must be in host
   * memory; every output buffer covers ALL cells.  Mutually exclusive with `comm` (one process per GPU). --- */
  int32_t  ngpu;           /* 0 / 1: single GPU                                                 */
  const int32_t* devices;  /* ngpu CUDA ordinals, or NULL: device, device + 1, ...              */
} scgbm_fit_args;

typedef struct scgbm_fit_out {
  /* caller-allocated, column-major doubles; V/beta/W cover this rank's cells */
  double* U;      /* I x M   last pgd_irlba left vectors, sign-fixed (R/fit.R:195,304-309) */
  double* D;      /* M                                                                       */
  double* V;      /* J x M   sign-fixed                                                      */
  double* loadings; /* I x M  U BEFORE the sign flip (R/fit.R:196, quirk Q7); may be NULL    */
  double* scores; /* J x M   V diag(D) (R/fit.R:310); may be NULL                            */
  double* alpha;  /* I x nbatch                                                              */
  double* beta;   /* J                                                                       */
  double* obj;    /* max_iter: accepted objectives (R/fit.R:174)                             */
  double* time;   /* max_iter or NULL: cumulative seconds (R/fit.R:139-142,203)              */
  double* W;      /* I x J or NULL (R/fit.R:189-191, unclipped)                              */
  double* LL;     /* max_iter or NULL: objective of EVERY iteration (diagnostic)             */
  int8_t* accepted; /* max_iter or NULL: 1 = accepted, 0 = reverted / break (diagnostic)     */
  /* scalars written by the library */
  int32_t n_obj;        /* entries of obj                                                    */
  int32_t n_iter;       /* loop trips, accepted or not                                       */
  int32_t hit_max_iter; /* R/fit.R:182-185 warning                                           */
  int32_t y_storage;    /* format actually used                                              */
  double  max_Y;
  scgbm_profile prof;
  /* extension: the factors of the X that alpha, beta and W belong to (the REVERTED X on a
   * convergence break, quirk Q6), so that get.se can re-evaluate W on the device when return.W = FALSE.
   * Caller-allocated or NULL.  x_rank = M (Xu = u diag(d), Xv = v), 0 (X = 0), or -1 when X is still the
   * clamped initial point of R/fit.R:116-120, which has no factor form (then Xu/Xv are not written). */
  double* Xu;     /* I x M */
  double* Xv;     /* J x M */
  int32_t x_rank;
  /* diagnostics of the truncated SVD that stands in for irlba (R/fit.R:115,323): irlba iterates to maxit and
   * warns; here the cycle budget is fixed, so calls that returned with a Ritz residual estimate >= svd_tol
   * are counted (the wrappers turn a non-zero count into a warning) */
  int32_t svd_unconverged;
  int32_t underflow_reruns; /* iterations re-evaluated on the CUDA-core pass because exp(eta) could underflow
                             * in the reference's doubles (log(0) = -Inf branch, R/fit.R:152-157)            */
  double  svd_worst_est;
  int64_t exact_rows;       /* gene row sums re-summed in fp64 over all iterations because |x| in the row grew
                             * beyond what the 22-bit tensor-core eta resolves to the alpha tolerance (12)   */
} scgbm_fit_out;

int scgbm_fit(const scgbm_fit_args* args, scgbm_fit_out* out);

/* The same fit, driven trip by trip (R/fit.R:125-186 is one trip).  `begin` uploads Y
 * and runs the setup + initial SVD (R/fit.R:76-123); `step` runs up to n_steps loop
 * trips and sets *done when the loop has ended (break, or max.iter reached); `end`
 * assembles the outputs (R/fit.R:188-210) and frees the handle (out may be NULL to
 * just free).  An R shim uses this to call R_CheckUserInterrupt between trips. */
typedef struct scgbm_fit_handle scgbm_fit_handle;
int scgbm_fit_begin(const scgbm_fit_args* args, scgbm_fit_handle** handle);
int scgbm_fit_step(scgbm_fit_handle* handle, int32_t n_steps, int32_t* done);
int scgbm_fit_profile(scgbm_fit_handle* handle, scgbm_profile* prof); /* cumulative so far */
int scgbm_fit_set_profile(scgbm_fit_handle* handle, int32_t level);    /* 0 / 1 / 2 as scgbm_fit_args.profile */
int scgbm_fit_end(scgbm_fit_handle* handle, scgbm_fit_out* out);

/* ---- get.se(out, EPS)   R/uncertainty.R:14-35 --------------------------------- */
typedef struct scgbm_se_args {
  int64_t I, J;         /* J = this rank's cells                                    */
  int32_t M;
  const double* U;      /* I x M  (out$U)                                           */
  const double* scores; /* J x M  (out$scores = V diag(D))                          */
  const double* W;      /* I x J  (out$W) or NULL -> W = exp(alpha + beta + Xu Xv^T)*/
  /* used when W == NULL (extension; W of the reverted X, quirk Q6):              */
  const double* alpha;  /* I x nbatch */
  const double* beta;   /* J          */
  const int32_t* batch; /* J or NULL  */
  int32_t nbatch;
  const double* Xu;     /* I x Mx left factor of X (already scaled by d) */
  const double* Xv;     /* J x Mx right factor of X                      */
  int32_t Mx;
  double  EPS;          /* R default 0 */
  int32_t device;
  scgbm_comm* comm;
  int32_t profile;
  int32_t ngpu;            /* >= 2: cells split over ngpu GPUs inside the library (as in scgbm_fit_args) */
  const int32_t* devices;
} scgbm_se_args;

typedef struct scgbm_se_out {
  double* se_scores;   /* J x M */
  double* se_loadings; /* I x M */
  int64_t n_singular;  /* matrices that were not positive definite */
  scgbm_profile prof;
} scgbm_se_out;

int scgbm_get_se(const scgbm_se_args* args, scgbm_se_out* out);

/* ---- projection step of gbm.sc(subset = ...)   R/fit.R:229-270 ---------------- */
typedef struct scgbm_proj_args {
  const void* Y;        /* I_full x J counts, column-major (all genes)              */
  int32_t  y_dtype;
  int32_t  y_on_device;
  int64_t  I_full, J, ldY;
  const int64_t* ixs;   /* Ik 0-based gene rows kept by the subsample fit (R/fit.R:225) */
  int64_t  Ik;
  int32_t  M;
  const double* U;      /* Ik x M  loadings of the subsample fit (R/fit.R:229)      */
  const double* alpha;  /* Ik      offset (R/fit.R:234)                             */
  double   glm_tol;     /* fastglm default 1e-8 */
  int32_t  glm_maxit;   /* fastglm default 100  */
  int32_t  device;
  int32_t  profile;
  /* sparse input, used when Y == NULL (same meaning as in scgbm_fit_args): I_full x J dgCMatrix slots */
  const int32_t* csc_i;
  const int64_t* csc_p;
  const void*    csc_x;
  int32_t  csc_x_dtype;
  /* how the per-cell GLMs are solved (0 = AUTO):
   *   SCGBM_GLM_EXACT    one fp64 IRLS per cell exactly as fastglm runs it (start mu = y + 0.1, deviance stop);
   *   SCGBM_GLM_TCGEN05  all cells in lock step on the tensor cores: Newton from the null model, gradient and
   *                      Hessian from the fused pass's product rider, stop on the Newton decrement (the predicted
   *                      deviance change); coefficients agree with the exact path to ~1e-6, `iters` counts Newton
   *                      solves and may differ from fastglm's; cells that do not converge are finished by the exact
   *                      kernel.  Needs integer counts (u8 / u16 storage).
   * AUTO: tcgen05 for integer counts and J >= 4096, else exact. */
  int32_t  glm_method;
} scgbm_proj_args;
enum scgbm_glm_method { SCGBM_GLM_AUTO = 0, SCGBM_GLM_EXACT = 1, SCGBM_GLM_TCGEN05 = 2 };

typedef struct scgbm_proj_out {
  double* beta;       /* J      intercept coefficient (R/fit.R:270)  */
  double* scores;     /* J x M  (R/fit.R:269)                        */
  double* se_scores;  /* J x M  (R/fit.R:268)                        */
  int8_t* fail;       /* J: 1 = GLM failed, row left at zero (R/fit.R:251-261) */
  int32_t* iters;     /* J or NULL: IRLS iterations used             */
  scgbm_profile prof;
  double* rowsum_full; /* I_full or NULL: rowSums(Y) of ALL genes over this call's cells (R/fit.R:217), computed from
                        * the same upload so that the host never has to sweep the I x J matrix itself */
  int32_t method_used;   /* SCGBM_GLM_EXACT or SCGBM_GLM_TCGEN05                                     */
  int32_t tc_sweeps;     /* tcgen05: lock-step Newton sweeps run                                      */
  int64_t tc_stragglers; /* tcgen05: cells handed to the exact kernel                                 */
} scgbm_proj_out;

int scgbm_project(const scgbm_proj_args* args, scgbm_proj_out* out);

/* ---- synthetic inputs on device (simData recipe, R/simulate.R:24-48) ---------- */
typedef struct scgbm_sim_args {
  int64_t I, J, ld;     /* ld >= I: leading dimension of the output               */
  int64_t j_offset;     /* global index of this shard's first cell (seeding)       */
  int32_t d;            /* latent factors; 0 -> null model                         */
  const double* U;      /* I x d (orthonormal) or NULL when d == 0                 */
  const double* V;      /* J x d rows of this shard                                */
  const double* D;      /* d                                                       */
  const double* alpha;  /* I                                                       */
  const double* beta;   /* J                                                       */
  double   nb_theta;    /* <= 0: Poisson; > 0: negative binomial (simNullNB)       */
  uint64_t seed;
  int32_t  out_dtype;   /* SCGBM_U16 / SCGBM_U8 / SCGBM_F32 / SCGBM_I32            */
  int32_t  device;
} scgbm_sim_args;

/* writes counts into the DEVICE buffer `Y_dev` (I x J, leading dim ld); returns the
 * number of entries that saturated the output type in *n_saturated. */
int scgbm_sim_counts(const scgbm_sim_args* args, void* Y_dev, int64_t* n_saturated);

/* ---- simData / simNull / simNullNB entirely on the device (R/simulate.R:24-48, :51-57, :60-65) ----
 * The reference's only other exported compute function (NAMESPACE:10) and the generator of every benchmark
 * input.  Everything is drawn on the GPU from one counter-based stream (Philox4x32-10 keyed by `seed`, one
 * counter per drawn element, so shards and launch geometry do not matter):
 *   SCGBM_SIM_DATA     U (I x d), V (J_total x d): Haar on the Stiefel manifold = QR of a Gaussian with
 *                      positive diag(R) (Cholesky-QR on device; rstiefel::rustiefel [ext] draws from the same
 *                      distribution), signs fixed so that U[1, m] >= 0 (:28-33);
 *                      D = seq(K (sqrt I + sqrt J), sqrt I + sqrt J, length.out = d) (:27);
 *                      alpha_i ~ N(alpha_mean, 1), beta_j ~ N(beta_mean, 1) (:35-36);
 *                      Y_ij ~ Poisson(exp(alpha_i + beta_j + (U D V^T)_ij)) (:37-38); theta > 0 draws
 *                      negative-binomial counts of that mean instead (MASS::rnegbin parametrisation, :63).
 *   SCGBM_SIM_NULL     mu_i ~ U(1, 10), Y_ij ~ Poisson(mu_i)                      (:51-57)
 *   SCGBM_SIM_NULL_NB  mu_i ~ U(1, 10), Y_ij ~ NB(mu_i, theta)                    (:60-65)
 * A call generates the cells [j_offset, j_offset + J) of a J_total-cell matrix; the gene-side draws and D
 * depend on (seed, I, J_total, d) only, so the shards of several ranks fit together. */
enum scgbm_sim_model { SCGBM_SIM_DATA = 0, SCGBM_SIM_NULL = 1, SCGBM_SIM_NULL_NB = 2 };

typedef struct scgbm_simdata_args {
  int32_t  model;        /* scgbm_sim_model                                                 */
  int32_t  d;            /* latent factors (SCGBM_SIM_DATA)                                 */
  int64_t  I, J;         /* genes; cells generated by THIS call                             */
  int64_t  J_total;      /* cells of the whole matrix (0: = J)                              */
  int64_t  j_offset;     /* global index of this call's first cell                          */
  double   alpha_mean, beta_mean, K;   /* R defaults 0, 0, 2                                */
  double   theta;        /* simNullNB: R default 1;  SCGBM_SIM_DATA: > 0 -> NB counts       */
  uint64_t seed;
  int32_t  out_dtype;    /* element type of Y: SCGBM_U8 / U16 / F32 / I32                   */
  int32_t  y_on_device;  /* out->Y is a device pointer                                      */
  int64_t  ldY;          /* leading dimension of out->Y (>= I; pad rows are written as 0)   */
  int32_t  device;
} scgbm_simdata_args;

typedef struct scgbm_simdata_out {
  void*   Y;             /* I x J counts, column-major, leading dim ldY                     */
  /* host, caller-allocated, each may be NULL: */
  double* U;             /* I x d      val$U                                                */
  double* V;             /* J x d      val$V = V diag(D) (ground-truth scores), this call's cells */
  double* D;             /* d          val$D                                                */
  double* alpha;         /* I          val$alpha  (null models: log(mu_i))                  */
  double* beta;          /* J          val$beta   (null models: 0)                          */
  int64_t n_saturated;   /* entries clipped to the maximum of out_dtype                     */
} scgbm_simdata_out;

int scgbm_sim_data(const scgbm_simdata_args* args, scgbm_simdata_out* out);

/* ---- CCI perturbation sampling (R/cci.R:69-71; the null draw of :57,:104) ---------------------------
 *   out[, , r] = scores + e_r,   e_r[j, m] ~ N(0, se_scores[j, m]^2),   r = 0 .. reps-1
 * -- the `V.new <- V + matrix(rnorm(J*M, sd = se_V), J, M)` that CCI() feeds to the user's clustering
 * function `reps` times -- as ONE batched device op.  scores == NULL gives e_r alone (V.null, R/cci.R:57).
 * Draw r of the stream is (seed, rep_offset + r), so R can fetch the replicates in batches. */
typedef struct scgbm_cci_args {
  int64_t  J;
  int32_t  M;
  const double* scores;     /* J x M (gbm$scores) or NULL                                   */
  const double* se_scores;  /* J x M (gbm$se_scores)                                        */
  int32_t  reps;
  int32_t  rep_offset;
  uint64_t seed;
  int32_t  out_on_device;   /* `out` is a device pointer                                    */
  int32_t  device;
} scgbm_cci_args;

int scgbm_cci_perturb(const scgbm_cci_args* args, double* out /* J x M x reps, column-major */);

/* ---- device memory helpers for callers that keep Y resident ------------------- */
int scgbm_dev_alloc(void** p, int64_t bytes, int device);
int scgbm_dev_free(void* p, int device);
int scgbm_memcpy_d2h(void* dst_host, const void* src_dev, int64_t bytes, int device);
int scgbm_memcpy_h2d(void* dst_dev, const void* src_host, int64_t bytes, int device);
int scgbm_host_alloc_pinned(void** p, int64_t bytes);
int scgbm_host_free_pinned(void* p);
int scgbm_flush_l2(int device);   /* writes a buffer larger than L2 */

/* ---- multi-GPU: one rank per GPU (NCCL loaded lazily with dlopen) ------------- */
#define SCGBM_UNIQUE_ID_BYTES 128
int  scgbm_comm_unique_id(uint8_t id[SCGBM_UNIQUE_ID_BYTES]);
int  scgbm_comm_init_rank(scgbm_comm** comm, const uint8_t id[SCGBM_UNIQUE_ID_BYTES],
                          int nranks, int rank, int device);
int  scgbm_comm_destroy(scgbm_comm* comm);
int  scgbm_comm_allreduce_f64(scgbm_comm* comm, double* host_buf, int64_t n, int op_max);

/* ---- misc ---------------------------------------------------------------------- */
const char* scgbm_last_error(void);
int  scgbm_device_count(void);
int  scgbm_version(void);

/* ---- diagnostic entry points (kernel-level parity tests and micro-benchmarks) -- */
/* truncated SVD of G = sum_t coef[t] * A_t diag(d_t) B_t^T + c * R  by the same
 * warm-started block-Krylov driver the fit uses.  R: I x J fp32 host, column-major. */
typedef struct scgbm_tsvd_args {
  int64_t I, J;
  const float* R;       /* I x J or NULL            */
  double  c;
  int32_t nterms;       /* 0..2                     */
  const double* A[2];   /* I x Mt[t]                */
  const double* d[2];   /* Mt[t]                    */
  const double* B[2];   /* J x Mt[t]                */
  int32_t Mt[2];
  double  coef[2];
  int32_t M;
  int32_t svd_extra, svd_depth;
  double  svd_tol;
  const double* V0;     /* J x (M+svd_extra) warm start or NULL */
  uint64_t seed;
  int32_t device;
  int32_t kernel_variant; /* 0 default, 1 force FFMA products, 2 force tcgen05 */
} scgbm_tsvd_args;

typedef struct scgbm_tsvd_out {
  double* U;  /* I x M */
  double* D;  /* M     */
  double* V;  /* J x M */
  int32_t passes;
  double  est;
} scgbm_tsvd_out;

int scgbm_dbg_tsvd(const scgbm_tsvd_args* args, scgbm_tsvd_out* out);

/* the two tall-skinny residual products of K5 on a given residual:
 *   outN (I x b) = c * R P,   outT (J x b) = c * R^T Q
 * R: I x J fp32 host (stored on device as bf16 hi/lo planes like the fit does);
 * variant: 1 = CUDA-core kernels, 2 = tcgen05 kernels.  ms_n / ms_t: CUDA-event time of
 * `reps` back-to-back launches of each product (mean per launch). */
typedef struct scgbm_prod_args {
  int64_t I, J;
  const float* R;
  const double* P;   /* J x b */
  const double* Q;   /* I x b */
  int32_t b;
  double  c;
  int32_t variant;
  int32_t reps;
  int32_t device;
} scgbm_prod_args;

typedef struct scgbm_prod_out {
  double* outN;      /* I x b */
  double* outT;      /* J x b */
  double* Rq;        /* I x J or NULL: hi + lo as stored (the matrix the products really use) */
  double  ms_n, ms_t;
} scgbm_prod_out;

int scgbm_dbg_prod(const scgbm_prod_args* args, scgbm_prod_out* out);

/* one evaluation of the loop body up to the objective (R/fit.R:127-151) for given
 * X = Xu Xv^T (optionally clamped to +-8), beta in / alpha,beta out. */
typedef struct scgbm_eval_args {
  const void* Y; int32_t y_dtype; int64_t I, J, ldY;
  int32_t y_storage;
  const double* Xu; const double* Xv; int32_t Mx; int32_t clamp8;
  const double* rowscale; /* I or NULL: X0 row scaling s_i   (R/fit.R:117) */
  const double* colscale; /* J or NULL: X0 col scaling t_j                  */
  const double* beta_in;  /* J */
  int32_t infer_beta;
  int32_t device;
} scgbm_eval_args;

typedef struct scgbm_eval_out {
  double* alpha;   /* I */
  double* beta;    /* J */
  double* resid;   /* I x J or NULL: Y - min(W, max.Y) as stored on device (fp32 -> double) */
  double  sum_w, sum_ylogw, max_w, max_Y, LL;
} scgbm_eval_out;

int scgbm_dbg_eval(const scgbm_eval_args* args, scgbm_eval_out* out);

/* symmetric eigen-decomposition on device (Jacobi): A (n x n, col-major) ->
 * eigenvalues descending in w, eigenvectors in the columns of A. */
int scgbm_dbg_eigh(double* A, double* w, int32_t n, int32_t device);

/* streaming rates of this GPU for two read : write mixes, plain coalesced kernels over 1 GiB streams:
 * gbs[0] = copy (1 read : 1 write, what MEASURED_PEAKS.json's hbm_gbs is), gbs[1] = 1 read : 2 writes (the traffic
 * pattern of the fused likelihood pass: Y in, two residual planes out).  bench.py reports both next to the roofline. */
int scgbm_dbg_membench(int32_t device, double gbs[2]);

#ifdef __cplusplus
}
#endif
#endif /* SCGBM_H */
