# The sampling step of CCI() (/root/reference/R/cci.R:69-71): `reps` perturbed score matrices
# V.new = gbm$scores + rnorm(J*M, sd = gbm$se_scores), drawn in one batched GPU call and returned as a
# J x M x reps array; feed [, , r] to the clustering function.  null = TRUE gives the perturbation alone
# (V.null, R/cci.R:57).  The rest of CCI (the user's clustering function, the overlap tables) stays in R.
cci.perturb <- function(gbm, reps = 100, seed = NULL, rep.offset = 0L, null = FALSE, device = -1L) {
  if (is.null(gbm$se_scores)) stop("run get.se() first: gbm$se_scores is missing")
  .Call("scgbm_cci_perturb_R", if (null) NULL else gbm$scores, gbm$se_scores, as.integer(reps),
        as.double(sim.seed(seed)), as.integer(rep.offset), as.integer(device), PACKAGE = "scGBM")
}
