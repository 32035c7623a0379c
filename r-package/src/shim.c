/* shim.c -- the R `.Call` boundary of libscgbm: pure marshalling between R objects and the plain-C structs of
 * include/scgbm.h.  No numerics happen here.  Replaces the bodies of
 *   gbm.sc            /root/reference/R/fit.R:52-211      -> scgbm_fit_R
 *   gbm.proj.parallel /root/reference/R/fit.R:236-266     -> scgbm_project_R   (the mclapply(fastglm) loop only)
 *   get.se            /root/reference/R/uncertainty.R:14-35 -> scgbm_get_se_R
 *   simData / simNull / simNullNB /root/reference/R/simulate.R:24-65 -> scgbm_sim_R
 *   the sampling step of CCI() /root/reference/R/cci.R:57,69-71       -> scgbm_cci_perturb_R
 * R is absent from the build image; tests/test_r_shim.py compiles and RUNS this file against the minimal
 * R-API stand-in under tests/r_stub/, so it cannot rot. */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "scgbm.h"

static void progress_cb(int iter, double obj, void* user) {      /* cat("Iteration: ", i, ". Objective=", LL[i]), R/fit.R:175 */
  (void)user;
  Rprintf("Iteration:  %d . Objective= %.7g\n", iter, obj);
}

static void set_elt(SEXP list, SEXP names, int k, const char* nm, SEXP v) {
  SET_VECTOR_ELT(list, k, v);
  SET_STRING_ELT(names, k, mkChar(nm));
}

static void check_interrupt(void* unused) { (void)unused; R_CheckUserInterrupt(); }

/* dense matrix (integer / double) or Matrix::dgCMatrix -> the count-matrix fields shared by fit and projection */
typedef struct { const void* Y; int dtype; R_xlen_t I, J; const int* ci; int64_t* cp; const double* cx; } counts_in;

static counts_in read_counts(SEXP Y) {
  counts_in c;
  memset(&c, 0, sizeof c);
  if (inherits(Y, "dgCMatrix")) {                                 /* sparse input: handed over column-compressed */
    SEXP dim = R_do_slot(Y, install("Dim"));
    SEXP sp = R_do_slot(Y, install("p"));
    c.I = INTEGER(dim)[0]; c.J = INTEGER(dim)[1];
    c.ci = INTEGER(R_do_slot(Y, install("i")));
    c.cx = REAL(R_do_slot(Y, install("x")));
    c.cp = (int64_t*)R_alloc((size_t)c.J + 1, sizeof(int64_t));   /* @p is 32-bit in Matrix, the ABI takes 64-bit */
    for (R_xlen_t j = 0; j <= c.J; ++j) c.cp[j] = INTEGER(sp)[j];
    c.dtype = SCGBM_F64;
    return c;
  }
  if (!isMatrix(Y) || !(isInteger(Y) || isReal(Y))) error("Y must be a numeric matrix or a dgCMatrix");
  c.I = nrows(Y); c.J = ncols(Y);
  c.Y = isInteger(Y) ? (const void*)INTEGER(Y) : (const void*)REAL(Y);
  c.dtype = isInteger(Y) ? SCGBM_I32 : SCGBM_F64;
  return c;
}

/* .Call("scgbm_fit_R", Y, M, max.iter, tol, min.iter, infer.beta, batch (integer codes or NULL), nbatch,
 *       return.W, time.by.iter, device, ngpu, keep.factors)
 * -> the list of R/fit.R:188-204,310 in the reference's element order (W, V, D, U, loadings, alpha, beta, M,
 *    [I, J,] obj, [time,] scores), plus, when keep.factors, Xu, Xv, x.rank for get.se after return.W = FALSE. */
SEXP scgbm_fit_R(SEXP Y, SEXP sM, SEXP sMaxIter, SEXP sTol, SEXP sMinIter, SEXP sInferBeta, SEXP sBatch,
                 SEXP sNbatch, SEXP sReturnW, SEXP sTime, SEXP sDevice, SEXP sNgpu, SEXP sKeep) {
  const counts_in c = read_counts(Y);
  const R_xlen_t I = c.I, J = c.J;
  const int M = asInteger(sM), max_iter = asInteger(sMaxIter), nbatch = asInteger(sNbatch);
  const int return_W = asLogical(sReturnW), time_by_iter = asLogical(sTime), keep = asLogical(sKeep);
  const int ngpu = asInteger(sNgpu);
  if (M < 1) error("M must be >= 1");                             /* R/fit.R:296-298 */
  if (max_iter < 1 || nbatch < 1) error("max.iter and the number of batches must be >= 1");
  if (!isNull(sBatch) && XLENGTH(sBatch) != J) error("Batch must be encoded as factor.");

  scgbm_fit_args a;
  memset(&a, 0, sizeof a);
  a.Y = c.Y; a.y_dtype = c.dtype; a.I = I; a.J = J; a.ldY = I;
  a.csc_i = c.ci; a.csc_p = c.cp; a.csc_x = c.cx; a.csc_x_dtype = SCGBM_F64;
  a.M = M; a.max_iter = max_iter; a.tol = asReal(sTol); a.min_iter = asInteger(sMinIter);
  a.infer_beta = asLogical(sInferBeta);
  a.batch = isNull(sBatch) ? NULL : INTEGER(sBatch); a.nbatch = nbatch;
  a.return_W = return_W; a.time_by_iter = time_by_iter; a.device = asInteger(sDevice);
  a.ngpu = ngpu;
  a.progress = progress_cb;

  int np = 0;
  SEXP U = PROTECT(allocMatrix(REALSXP, (int)I, M)); np++;
  SEXP V = PROTECT(allocMatrix(REALSXP, (int)J, M)); np++;
  SEXP D = PROTECT(allocVector(REALSXP, M)); np++;
  SEXP loadings = PROTECT(allocMatrix(REALSXP, (int)I, M)); np++;          /* U before the sign flip: R/fit.R:196 */
  SEXP scores = PROTECT(allocMatrix(REALSXP, (int)J, M)); np++;
  SEXP alpha = PROTECT(allocMatrix(REALSXP, (int)I, nbatch)); np++;
  SEXP beta = PROTECT(allocVector(REALSXP, J)); np++;
  SEXP obj = PROTECT(allocVector(REALSXP, max_iter)); np++;
  SEXP tm = PROTECT(allocVector(REALSXP, max_iter)); np++;
  SEXP W = R_NilValue, Xu = R_NilValue, Xv = R_NilValue;
  if (return_W) { W = PROTECT(allocMatrix(REALSXP, (int)I, (int)J)); np++; }
  if (keep) { Xu = PROTECT(allocMatrix(REALSXP, (int)I, M)); np++; Xv = PROTECT(allocMatrix(REALSXP, (int)J, M)); np++; }

  scgbm_fit_out o;
  memset(&o, 0, sizeof o);
  o.U = REAL(U); o.D = REAL(D); o.V = REAL(V); o.loadings = REAL(loadings); o.scores = REAL(scores);
  o.alpha = REAL(alpha); o.beta = REAL(beta); o.obj = REAL(obj);
  o.time = time_by_iter ? REAL(tm) : NULL;
  o.W = return_W ? REAL(W) : NULL;
  if (keep) { o.Xu = REAL(Xu); o.Xv = REAL(Xv); }

  int st;
  if (ngpu > 1) {
    st = scgbm_fit(&a, &o);            /* several GPUs from this one call: host threads live inside the library */
  } else {
    /* trip by trip (one trip = R/fit.R:125-186) so that Ctrl-C is honoured between trips; R_CheckUserInterrupt
     * longjmps, so it runs under R_ToplevelExec and the device handle is released before the error is raised */
    scgbm_fit_handle* h = NULL;
    int done = 0;
    st = scgbm_fit_begin(&a, &h);
    while (st == 0 && !done) {
      st = scgbm_fit_step(h, 1, &done);
      if (st == 0 && !done && R_ToplevelExec(check_interrupt, NULL) == FALSE) {
        scgbm_fit_end(h, NULL);
        UNPROTECT(np);
        error("interrupted");
      }
    }
    if (st == 0) st = scgbm_fit_end(h, &o);
    else if (h) scgbm_fit_end(h, NULL);
  }
  if (st != 0) { UNPROTECT(np); error("%s", scgbm_last_error()); }         /* the library has freed its device memory */
  if (o.hit_max_iter)                                                      /* R/fit.R:182-185 */
    warning("Maximum number of iterations reached (increase max.iter).\n              Possible non-convergence.");
  if (o.svd_unconverged)
    warning("truncated SVD: %d call(s) stopped above the residual tolerance (worst estimate %.2e)", o.svd_unconverged, o.svd_worst_est);

  /* V, D, U, loadings, alpha, beta, M, obj, scores = 9; W, I, J only with W; time; Xu, Xv, x.rank */
  const int n = (return_W ? 3 : 0) + 9 + (time_by_iter ? 1 : 0) + (keep ? 3 : 0);
  SEXP out = PROTECT(allocVector(VECSXP, n)); np++;
  SEXP nm = PROTECT(allocVector(STRSXP, n)); np++;
  int k = 0;
  if (return_W) set_elt(out, nm, k++, "W", W);
  set_elt(out, nm, k++, "V", V);
  set_elt(out, nm, k++, "D", D);
  set_elt(out, nm, k++, "U", U);
  set_elt(out, nm, k++, "loadings", loadings);
  set_elt(out, nm, k++, "alpha", alpha);
  set_elt(out, nm, k++, "beta", beta);
  set_elt(out, nm, k++, "M", ScalarInteger(M));
  if (return_W) {                                                 /* I <- nrow(W): absent when W is (R/fit.R:199-200) */
    set_elt(out, nm, k++, "I", ScalarInteger((int)I));
    set_elt(out, nm, k++, "J", ScalarInteger((int)J));
  }
  set_elt(out, nm, k++, "obj", lengthgets(obj, o.n_obj));
  if (time_by_iter) set_elt(out, nm, k++, "time", lengthgets(tm, o.n_iter));
  set_elt(out, nm, k++, "scores", scores);
  if (keep) {
    set_elt(out, nm, k++, "Xu", Xu);
    set_elt(out, nm, k++, "Xv", Xv);
    set_elt(out, nm, k++, "x.rank", ScalarInteger(o.x_rank));
  }
  setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(np);
  return out;
}

/* .Call("scgbm_get_se_R", U, scores, W (matrix or NULL), alpha, beta, batch, Xu, Xv, x.rank, EPS, device, ngpu)
 * -> list(se_scores, se_loadings).  W = NULL (a return.W = FALSE fit): the reference stops there
 * (R/uncertainty.R:18 reads out$W); with the kept factors W is re-evaluated on the device instead. */
SEXP scgbm_get_se_R(SEXP U, SEXP S, SEXP W, SEXP sAlpha, SEXP sBeta, SEXP sBatch, SEXP sXu, SEXP sXv, SEXP sXrank,
                    SEXP sEPS, SEXP sDevice, SEXP sNgpu) {
  if (!isMatrix(U) || !isMatrix(S)) error("out$U and out$scores must be matrices");
  scgbm_se_args a;
  memset(&a, 0, sizeof a);
  a.I = nrows(U); a.J = nrows(S); a.M = ncols(U);
  a.U = REAL(U); a.scores = REAL(S);
  a.EPS = asReal(sEPS); a.device = asInteger(sDevice); a.ngpu = asInteger(sNgpu); a.nbatch = 1;
  if (!isNull(W)) {
    if (!isMatrix(W) || nrows(W) != a.I || ncols(W) != a.J) error("out$W does not match out$U / out$scores");
    a.W = REAL(W);
  } else {
    if (isNull(sXu) || isNull(sXv) || isNull(sAlpha) || isNull(sBeta) || asInteger(sXrank) < 0)
      error("get.se needs out$W: fit with return.W = TRUE, or with keep.factors = TRUE to evaluate W on the GPU");
    a.alpha = REAL(sAlpha); a.beta = REAL(sBeta);
    a.nbatch = isMatrix(sAlpha) ? ncols(sAlpha) : 1;
    a.batch = isNull(sBatch) ? NULL : INTEGER(sBatch);
    a.Xu = REAL(sXu); a.Xv = REAL(sXv); a.Mx = asInteger(sXrank);
  }
  SEXP se_s = PROTECT(allocMatrix(REALSXP, (int)a.J, a.M));
  SEXP se_l = PROTECT(allocMatrix(REALSXP, (int)a.I, a.M));
  scgbm_se_out o;
  memset(&o, 0, sizeof o);
  o.se_scores = REAL(se_s); o.se_loadings = REAL(se_l);
  if (scgbm_get_se(&a, &o) != 0) { UNPROTECT(2); error("%s", scgbm_last_error()); }   /* solve() would stop() too */
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, se_s);
  SET_VECTOR_ELT(out, 1, se_l);
  UNPROTECT(3);
  return out;
}

/* .Call("scgbm_project_R", Y, ixs (1-based kept genes), U, alpha, device) -> list(beta, scores, se_scores, fail) */
SEXP scgbm_project_R(SEXP Y, SEXP sIxs, SEXP U, SEXP sAlpha, SEXP sDevice) {
  const counts_in c = read_counts(Y);
  const R_xlen_t Ik = XLENGTH(sIxs);
  if (!isMatrix(U) || nrows(U) != Ik || XLENGTH(sAlpha) != Ik) error("U, alpha and ixs disagree");
  const int M = ncols(U);
  int64_t* ixs = (int64_t*)R_alloc((size_t)Ik, sizeof(int64_t));
  for (R_xlen_t k = 0; k < Ik; ++k) ixs[k] = (int64_t)INTEGER(sIxs)[k] - 1;          /* which() is 1-based */
  scgbm_proj_args a;
  memset(&a, 0, sizeof a);
  a.Y = c.Y; a.y_dtype = c.dtype; a.I_full = c.I; a.J = c.J; a.ldY = c.I;
  a.csc_i = c.ci; a.csc_p = c.cp; a.csc_x = c.cx; a.csc_x_dtype = SCGBM_F64;
  a.ixs = ixs; a.Ik = Ik; a.M = M; a.U = REAL(U); a.alpha = REAL(sAlpha);
  a.glm_tol = 1e-8; a.glm_maxit = 100;                                               /* fastglm's defaults */
  a.device = asInteger(sDevice);
  SEXP beta = PROTECT(allocVector(REALSXP, c.J));
  SEXP sc = PROTECT(allocMatrix(REALSXP, (int)c.J, M));
  SEXP se = PROTECT(allocMatrix(REALSXP, (int)c.J, M));
  SEXP fail = PROTECT(allocVector(RAWSXP, c.J));
  scgbm_proj_out o;
  memset(&o, 0, sizeof o);
  o.beta = REAL(beta); o.scores = REAL(sc); o.se_scores = REAL(se); o.fail = (int8_t*)RAW(fail);
  if (scgbm_project(&a, &o) != 0) { UNPROTECT(4); error("%s", scgbm_last_error()); }
  SEXP out = PROTECT(allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, beta); SET_VECTOR_ELT(out, 1, sc); SET_VECTOR_ELT(out, 2, se); SET_VECTOR_ELT(out, 3, fail);
  UNPROTECT(5);
  return out;
}

/* .Call("scgbm_sim_R", model (0 simData, 1 simNull, 2 simNullNB), I, J, d, alpha.mean, beta.mean, K, theta, seed, device)
 * -> list(Y, V, U, D, alpha, beta) as in R/simulate.R:40-46 (null models: Y and the gene means) */
SEXP scgbm_sim_R(SEXP sModel, SEXP sI, SEXP sJ, SEXP sd, SEXP sAm, SEXP sBm, SEXP sK, SEXP sTheta, SEXP sSeed, SEXP sDevice) {
  const int model = asInteger(sModel), I = asInteger(sI), J = asInteger(sJ), d = model == SCGBM_SIM_DATA ? asInteger(sd) : 0;
  if (I < 1 || J < 1) error("I and J must be positive");
  scgbm_simdata_args a;
  memset(&a, 0, sizeof a);
  a.model = model; a.d = d; a.I = I; a.J = J; a.J_total = J;
  a.alpha_mean = asReal(sAm); a.beta_mean = asReal(sBm); a.K = asReal(sK); a.theta = asReal(sTheta);
  a.seed = (uint64_t)asReal(sSeed); a.out_dtype = SCGBM_I32; a.ldY = I; a.device = asInteger(sDevice);
  SEXP Y = PROTECT(allocMatrix(INTSXP, I, J));
  SEXP U = PROTECT(allocMatrix(REALSXP, I, d > 0 ? d : 1));
  SEXP V = PROTECT(allocMatrix(REALSXP, J, d > 0 ? d : 1));
  SEXP D = PROTECT(allocVector(REALSXP, d > 0 ? d : 1));
  SEXP alpha = PROTECT(allocVector(REALSXP, I));
  SEXP beta = PROTECT(allocVector(REALSXP, J));
  scgbm_simdata_out o;
  memset(&o, 0, sizeof o);
  o.Y = INTEGER(Y); o.alpha = REAL(alpha); o.beta = REAL(beta);
  if (d > 0) { o.U = REAL(U); o.V = REAL(V); o.D = REAL(D); }
  if (scgbm_sim_data(&a, &o) != 0) { UNPROTECT(6); error("%s", scgbm_last_error()); }
  SEXP out = PROTECT(allocVector(VECSXP, 6));
  SEXP nm = PROTECT(allocVector(STRSXP, 6));
  set_elt(out, nm, 0, "Y", Y); set_elt(out, nm, 1, "V", V); set_elt(out, nm, 2, "U", U);
  set_elt(out, nm, 3, "D", D); set_elt(out, nm, 4, "alpha", alpha); set_elt(out, nm, 5, "beta", beta);
  setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(8);
  return out;
}

/* .Call("scgbm_cci_perturb_R", scores (or NULL for the null draw), se_scores, reps, seed, rep.offset, device)
 * -> a J x M x reps array: [, , r] = V + rnorm(J*M, sd = se_V)   (R/cci.R:69-71; R/cci.R:57 when scores is NULL) */
SEXP scgbm_cci_perturb_R(SEXP S, SEXP SE, SEXP sReps, SEXP sSeed, SEXP sOffset, SEXP sDevice) {
  if (!isMatrix(SE)) error("gbm$se_scores must be a matrix (run get.se first)");
  const int J = nrows(SE), M = ncols(SE), reps = asInteger(sReps);
  if (!isNull(S) && (!isMatrix(S) || nrows(S) != J || ncols(S) != M)) error("scores and se_scores differ in shape");
  if (reps < 1) error("reps must be >= 1");
  scgbm_cci_args a;
  memset(&a, 0, sizeof a);
  a.J = J; a.M = M; a.scores = isNull(S) ? NULL : REAL(S); a.se_scores = REAL(SE);
  a.reps = reps; a.rep_offset = asInteger(sOffset); a.seed = (uint64_t)asReal(sSeed); a.device = asInteger(sDevice);
  SEXP out = PROTECT(alloc3DArray(REALSXP, J, M, reps));
  if (scgbm_cci_perturb(&a, REAL(out)) != 0) { UNPROTECT(1); error("%s", scgbm_last_error()); }
  UNPROTECT(1);
  return out;
}

SEXP scgbm_device_count_R(void) { return ScalarInteger(scgbm_device_count()); }

static const R_CallMethodDef calls[] = {
  {"scgbm_fit_R", (DL_FUNC)&scgbm_fit_R, 13},
  {"scgbm_get_se_R", (DL_FUNC)&scgbm_get_se_R, 12},
  {"scgbm_project_R", (DL_FUNC)&scgbm_project_R, 5},
  {"scgbm_sim_R", (DL_FUNC)&scgbm_sim_R, 10},
  {"scgbm_cci_perturb_R", (DL_FUNC)&scgbm_cci_perturb_R, 6},
  {"scgbm_device_count_R", (DL_FUNC)&scgbm_device_count_R, 0},
  {NULL, NULL, 0}};

void R_init_scGBM(DllInfo* dll) {
  R_registerRoutines(dll, NULL, calls, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
