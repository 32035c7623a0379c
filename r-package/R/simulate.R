# simData / simNull / simNullNB drawn on the GPU (/root/reference/R/simulate.R:24-65).  The draws follow the
# reference's distributions from the library's own counter-based stream; `seed` defaults to a draw from R's RNG,
# so set.seed() still makes a script reproducible.
sim.seed <- function(seed) if (is.null(seed)) sample.int(.Machine$integer.max, 1) else seed

simData <- function(I, J, d, alpha.mean = 0, beta.mean = 0, K = 2, seed = NULL, device = -1L) {
  val <- .Call("scgbm_sim_R", 0L, as.integer(I), as.integer(J), as.integer(d), as.double(alpha.mean),
               as.double(beta.mean), as.double(K), 0, as.double(sim.seed(seed)), as.integer(device), PACKAGE = "scGBM")
  val                                     # list(Y, V = scores, U, D, alpha, beta), R/simulate.R:40-46
}

simNull <- function(I, J, seed = NULL, device = -1L) {
  .Call("scgbm_sim_R", 1L, as.integer(I), as.integer(J), 0L, 0, 0, 2, 0, as.double(sim.seed(seed)),
        as.integer(device), PACKAGE = "scGBM")$Y
}

simNullNB <- function(I, J, theta = 1, seed = NULL, device = -1L) {
  .Call("scgbm_sim_R", 2L, as.integer(I), as.integer(J), 0L, 0, 0, 2, as.double(theta), as.double(sim.seed(seed)),
        as.integer(device), PACKAGE = "scGBM")$Y
}

scgbm.device.count <- function() .Call("scgbm_device_count_R", PACKAGE = "scGBM")
