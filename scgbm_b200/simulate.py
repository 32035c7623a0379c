"""Host-side mirror of the reference's generators and of CCI's perturbation step, on top of the C ABI:

    simData(I, J, d, alpha.mean, beta.mean, K)     R/simulate.R:24-48
    simNull(I, J)                                  R/simulate.R:51-57
    simNullNB(I, J, theta)                         R/simulate.R:60-65
    V + rnorm(J*M, sd = se_V)  x  reps             R/cci.R:69-71 (and the null draw, :57)

Everything is drawn on the GPU (scgbm_sim_data, scgbm_cci_perturb) from a counter-based Philox stream keyed by
`seed` -- R's own RNG stream is not restated (the draws are the reference's distributions, not its numbers)."""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib as L

_DT = {np.dtype(np.int32): L.I32, np.dtype(np.uint16): L.U16, np.dtype(np.uint8): L.U8, np.dtype(np.float32): L.F32}


def _sim(model, I, J, d, alpha_mean, beta_mean, K, theta, seed, dtype, device, j_offset=0, J_total=0, want_factors=True):
    lib = L.load()
    dtype = np.dtype(dtype)
    if dtype not in _DT:
        raise ValueError("dtype must be int32, uint16, uint8 or float32")
    I, J, d = int(I), int(J), int(d)
    Y = np.zeros((I, J), dtype=dtype, order="F")
    a = L.SimDataArgs()
    a.model = model; a.d = d; a.I = I; a.J = J; a.J_total = int(J_total); a.j_offset = int(j_offset)
    a.alpha_mean = float(alpha_mean); a.beta_mean = float(beta_mean); a.K = float(K); a.theta = float(theta)
    a.seed = int(seed); a.out_dtype = _DT[dtype]; a.y_on_device = 0; a.ldY = I; a.device = int(device)
    o = L.SimDataOut()
    o.Y = L.ptr(Y)
    U = V = D = None
    alpha = np.zeros(I); beta = np.zeros(J)
    o.alpha = L.ptr(alpha); o.beta = L.ptr(beta)
    if want_factors and d > 0:
        U = np.zeros((I, d), order="F"); V = np.zeros((J, d), order="F"); D = np.zeros(d)
        o.U = L.ptr(U); o.V = L.ptr(V); o.D = L.ptr(D)
    L.check(lib.scgbm_sim_data(C.byref(a), C.byref(o)))
    return Y, U, V, D, alpha, beta, int(o.n_saturated)


def sim_data(I, J, d, alpha_mean=0.0, beta_mean=0.0, K=2.0, *, seed=0, theta=0.0, dtype=np.int32, device=-1,
             j_offset=0, J_total=0):
    """simData(I, J, d, alpha.mean = 0, beta.mean = 0, K = 2)  -- R/simulate.R:24-48.  Returns the reference's list
    (Y, V = scores, U, D, alpha, beta).  Extensions: `seed`; `theta` > 0 draws negative-binomial counts of the same
    mean (the benchmark configs); `j_offset`/`J_total` generate one shard of a larger matrix."""
    Y, U, V, D, alpha, beta, nsat = _sim(0, I, J, d, alpha_mean, beta_mean, K, theta, seed, dtype, device, j_offset, J_total)
    val = OrderedDict()
    val["Y"] = Y; val["V"] = V; val["U"] = U; val["D"] = D; val["alpha"] = alpha; val["beta"] = beta   # :40-46
    val["_n_saturated"] = nsat
    return val


def sim_null(I, J, *, seed=0, dtype=np.int32, device=-1, return_mu=False):
    """simNull(I, J)  -- R/simulate.R:51-57: mu_i ~ U(1, 10), Y_ij ~ Poisson(mu_i).  Returns Y like the reference."""
    Y, _, _, _, alpha, _, _ = _sim(1, I, J, 0, 0.0, 0.0, 2.0, 0.0, seed, dtype, device, want_factors=False)
    return (Y, np.exp(alpha)) if return_mu else Y


def sim_null_nb(I, J, theta=1.0, *, seed=0, dtype=np.int32, device=-1, return_mu=False):
    """simNullNB(I, J, theta = 1)  -- R/simulate.R:60-65: mu_i ~ U(1, 10), Y_ij ~ NB(mu_i, theta) (MASS::rnegbin)."""
    Y, _, _, _, alpha, _, _ = _sim(2, I, J, 0, 0.0, 0.0, 2.0, theta, seed, dtype, device, want_factors=False)
    return (Y, np.exp(alpha)) if return_mu else Y


def cci_perturb(gbm, reps=100, *, seed=0, rep_offset=0, null=False, device=-1):
    """The sampling step of CCI(gbm, ..., reps)  -- R/cci.R:52-53,69-71: `reps` matrices V.new = gbm$scores + e with
    e[j, m] ~ N(0, gbm$se_scores[j, m]^2), as one (J, M, reps) array (Fortran order: [:, :, r] is replicate r).
    null=True gives e alone -- the V.null draw of R/cci.R:57.  Feed each slice to the clustering function."""
    lib = L.load()
    se = np.asfortranarray(gbm["se_scores"], dtype=np.float64)
    J, M = se.shape
    sc = None
    if not null:
        sc = np.asfortranarray(gbm["scores"], dtype=np.float64)
        if sc.shape != se.shape:
            raise ValueError("scores and se_scores must have the same shape")
    out = np.zeros((J, M, int(reps)), order="F")
    a = L.CciArgs()
    a.J = J; a.M = M; a.scores = L.ptr(sc); a.se_scores = L.ptr(se); a.reps = int(reps); a.rep_offset = int(rep_offset)
    a.seed = int(seed); a.out_on_device = 0; a.device = int(device)
    L.check(lib.scgbm_cci_perturb(C.byref(a), L.ptr(out)))
    return out
