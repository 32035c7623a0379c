# get.se() on the GPU: per-cell and per-gene information matrices, their inverses' diagonals
# (/root/reference/R/uncertainty.R:14-35).  Works from out$W like the reference, or -- after
# gbm.sc(..., return.W = FALSE, keep.factors = TRUE) -- from the kept factors, W being re-evaluated on the device.
get.se <- function(out, EPS = 0, device = -1L, ngpu = 1L) {
  if (is.null(out$W) && is.null(out$Xu))
    stop("get.se needs out$W: fit with return.W = TRUE, or with keep.factors = TRUE to evaluate W on the GPU")
  se <- .Call("scgbm_get_se_R", out$U, out$scores, out$W, out$alpha, out$beta, out$batch, out$Xu, out$Xv,
              if (is.null(out$x.rank)) -1L else as.integer(out$x.rank), as.double(EPS), as.integer(device),
              as.integer(ngpu), PACKAGE = "scGBM")
  out$se_scores <- se[[1]]                                                # R/uncertainty.R:26
  out$se_loadings <- se[[2]]                                              # R/uncertainty.R:33
  out
}
