# gbm.sc() on the GPU.  Same signature, defaults, conditions and returned list as the reference's
# /root/reference/R/fit.R:52-62,188-210; everything between argument checking and list decoration happens in
# libscgbm (src/shim.c -> include/scgbm.h).  Extra arguments, all optional and after the reference's own:
#   device        CUDA ordinal (-1: current)
#   ngpu          >= 2: this one call splits the cells over that many GPUs of the box
#   keep.factors  also return the factors of the X that alpha, beta and W belong to (Xu, Xv, x.rank), so that
#                 get.se() works after return.W = FALSE (in the reference it cannot: it reads out$W)

gbm.sc <- function(Y, M, max.iter = 100, tol = 10^{-4}, subset = NULL, ncores = 1, infer.beta = FALSE,
                   return.W = TRUE, batch = as.factor(rep(1, ncol(Y))), time.by.iter = FALSE, min.iter = 30,
                   device = -1L, ngpu = 1L, keep.factors = FALSE) {
  sparse <- methods::is(Y, "dgCMatrix")
  if (!sparse && !is.matrix(Y)) Y <- as.matrix(Y)
  check.counts(if (sparse) Y@x else Y, M)                     # the three stop()s of R/fit.R:287-299

  if (!is.null(subset)) {                                     # projection method, R/fit.R:67-73
    out <- gbm.proj.parallel(Y, M, subsample = subset, ncores = ncores, tol = tol, max.iter = max.iter,
                             device = device)
    message("For users of newer versions (1.0.1+): the `scores` matrix now contains factor scores.")
    return(out)
  }

  if (!is.factor(batch)) stop("Batch must be encoded as factor.")          # R/fit.R:86-88
  codes <- as.integer(batch)                                               # R/fit.R:90
  nbatch <- max(codes)                                                     # R/fit.R:91
  out <- .Call("scgbm_fit_R", Y, as.integer(M), as.integer(max.iter), as.double(tol), as.integer(min.iter),
               as.logical(infer.beta), if (nbatch > 1L) codes else NULL, as.integer(nbatch),
               as.logical(return.W), as.logical(time.by.iter), as.integer(device), as.integer(ngpu),
               as.logical(keep.factors), PACKAGE = "scGBM")

  # dimnames exactly where the reference puts them (R/fit.R:193-198)
  cells <- colnames(Y); genes <- rownames(Y)
  rownames(out$V) <- cells; colnames(out$V) <- 1:M
  names(out$D) <- 1:M
  rownames(out$U) <- genes; colnames(out$U) <- 1:M
  rownames(out$loadings) <- genes; colnames(out$loadings) <- 1:M
  rownames(out$alpha) <- genes
  names(out$beta) <- cells
  rownames(out$scores) <- cells; colnames(out$scores) <- 1:M
  if (keep.factors) out$batch <- if (nbatch > 1L) codes else NULL
  message("For users of newer versions (1.0.1+): the `scores` matrix now contains factor scores, the `V` matrix is UNSCALED scores.")
  out
}

# NA / negative / M < 1 (R/fit.R:287-299), on the stored values only when Y is sparse
check.counts <- function(y, M) {
  if (anyNA(y)) stop("Count matrix Y cannot contain NA values.")
  if (length(y) > 0 && min(y) < 0) stop("Count matrix Y cannot contain negative values.")
  if (M < 1) stop("M must be >= 1")
  invisible(NULL)
}

# gbm.sc(subset = ...): fit on a subsample of the cells, then one Poisson GLM per cell on the fitted loadings
# (R/fit.R:213-285).  The draw, the gene filter, the recursive fit and the re-centring stay in R -- so sample() keeps
# R's RNG semantics --; the J per-cell regressions that the reference farms out to mclapply(fastglm) in chunks of
# 200 000 cells (R/fit.R:236-266) are ONE call into the library.
gbm.proj.parallel <- function(Y, M, subsample = 2000, min.counts = 5, ncores = 1, tol = 10^{-4}, max.iter = 100,
                              device = -1L) {
  J <- ncol(Y); I <- nrow(Y)
  gene.total <- if (methods::is(Y, "dgCMatrix")) Matrix::rowSums(Y) else rowSums(Y)
  log.gene.total <- log(gene.total)                                       # R/fit.R:217
  jxs <- if (length(subsample) == 1) sample(1:J, size = subsample, replace = FALSE) else subsample   # :218-223
  Y.sub <- as.matrix(Y[, jxs])                                            # :224
  ixs <- which(rowSums(Y.sub) > 5)                                        # :225  (the literal 5 of the reference)
  out <- gbm.sc(Y.sub[ixs, , drop = FALSE], M = M, tol = tol, max.iter = max.iter, device = device)   # :227

  fit <- .Call("scgbm_project_R", Y, as.integer(ixs), out$U, as.double(out$alpha[, 1]), as.integer(device),
               PACKAGE = "scGBM")
  beta <- fit[[1]]
  out$scores <- fit[[2]]; colnames(out$scores) <- 1:M                     # :269
  out$se_scores <- fit[[3]]; colnames(out$se_scores) <- 1:M               # :268
  # cells whose regression failed keep an all-zero row, like the try({...}) of R/fit.R:251-261 (fit[[4]] flags them)

  alpha <- numeric(I)                                                     # :272-277
  alpha[ixs] <- out$alpha[, 1]
  if (length(ixs) < I) alpha[-ixs] <- log.gene.total[-ixs] - log(sum(exp(beta)))
  centre <- mean(alpha)
  out$beta <- beta + centre
  out$alpha <- alpha - centre
  U <- matrix(0, nrow = I, ncol = M)                                      # :278-280
  U[ixs, ] <- out$U
  out$U <- U
  out$V <- NULL                                                           # :281
  out
}
