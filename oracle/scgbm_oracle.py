"""TEST INFRASTRUCTURE ONLY -- float64 CPU restatement of the reference's hot path.

Restates, line for line, `/root/reference/R/fit.R` (`gbm.sc` :52-211,
`gbm.proj.parallel` :213-285, `gbm.sc.check.valid.input` :287-299,
`process.results` :301-313, `pgd_irlba` :315-328) and
`/root/reference/R/uncertainty.R` (`get.se` :14-35) in numpy.

R is not installed in this image, so the reference itself cannot be executed.
Third-party numerics the reference delegates to are replaced by mathematically
equivalent, tighter solvers:
  * irlba::irlba (un-pinned CRAN dep, R/fit.R:115,323; tol 1e-5 Lanczos) ->
    exact LAPACK SVD truncated to M (`svd="exact"`) or ARPACK svds
    (`svd="svds"`, the timing stand-in).
  * fastglm::fastglm(method=3) (un-declared CRAN dep, R/fit.R:253) -> IRLS with
    Cholesky, same start (mu = y + 0.1) and stopping rule (|ddev|/(|dev|+0.1) <
    1e-8, <= 100 its) as published for fastglm [ext].
  * base::solve (R/uncertainty.R:24,31) -> numpy.linalg.inv.

PINNING: the only golden vectors in the reference are the README demo's printed
objective trace (README.md:36-62) and se_scores head (README.md:101-114).  With
R's RNG restated in oracle/r_rng.py this oracle reproduces them
(tests/test_oracle_kat.py), which is what pins it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product never does.
"""
from __future__ import annotations

import math
import time as _time
import numpy as np
import scipy.linalg as sla


class RError(RuntimeError):
    """R `stop()` condition."""


# ----------------------------------------------------------------------------
# truncated SVD stand-ins for irlba::irlba(A, nv=M)   (R/fit.R:115, :323)
# ----------------------------------------------------------------------------
def _tsvd_exact(A: np.ndarray, M: int):
    u, d, vt = sla.svd(A, full_matrices=False, lapack_driver="gesdd", check_finite=False)
    return u[:, :M].copy(), d[:M].copy(), vt[:M, :].T.copy()


def _tsvd_svds(A: np.ndarray, M: int, rng=None):
    from scipy.sparse.linalg import svds
    v0 = None
    if rng is not None:
        v0 = rng.standard_normal(min(A.shape))
    u, d, vt = svds(A, k=M, tol=1e-5, v0=v0)
    o = np.argsort(-d)
    return u[:, o].copy(), d[o].copy(), vt[o, :].T.copy()


def _tsvd(A, M, svd, rng=None):
    if callable(svd):            # test hook: alternative truncated-SVD models
        return svd(A, M)
    if svd == "exact":
        return _tsvd_exact(A, M)
    if svd == "svds":
        return _tsvd_svds(A, M, rng)
    raise ValueError(svd)


# ----------------------------------------------------------------------------
def check_valid_input(Y, M):
    """R/fit.R:287-299."""
    if np.isnan(Y).any():
        raise RError("Count matrix Y cannot contain NA values.")
    if Y.min() < 0:
        raise RError("Count matrix Y cannot contain negative values.")
    if M < 1:
        raise RError("M must be >= 1")


def pgd_irlba(X, Xt, i, lr, W, Y, M, svd, rng=None):
    """R/fit.R:315-328.  `i` is the 1-based loop index."""
    V = X + ((i - 1) / (i + 2)) * (X - Xt)          # :319
    Xt_new = X                                       # :320
    w_max = W.max()                                  # :321
    u, d, v = _tsvd(V + (lr / w_max) * (Y - W), M, svd, rng)   # :323
    Xn = u @ (d[:, None] * v.T)                      # :324
    return Xn, Xt_new, (u, d, v)


def process_results(out):
    """R/fit.R:301-313: U[1,m] >= 0 sign convention, scores = V diag(D)."""
    U = out["U"].copy()
    V = out["V"].copy()
    for m in range(out["M"]):
        if U[0, m] < 0:
            U[:, m] = -U[:, m]
            V[:, m] = -V[:, m]
    out["U"] = U
    out["V"] = V
    out["scores"] = V * out["D"][None, :]
    return out


def gbm_sc(Y, M, max_iter=100, tol=1e-4, subset=None, ncores=1, infer_beta=False,
           return_W=True, batch=None, time_by_iter=False, min_iter=30,
           svd="exact", rng=None, verbose=False, _ll_inject=None):
    """R/fit.R:52-211.  `batch` is an int array of 1-based codes (as.integer of
    the factor, :90) or None for a single batch.  `_ll_inject` ({iteration: value}, test hook) replaces the
    objective of those iterations, to drive the NaN/Inf branch of :152-157 (quirks Q1, Q2) at will."""
    Y = np.asarray(Y, dtype=np.float64)
    check_valid_input(Y, M)                                            # :65
    if subset is not None:                                             # :67-73
        return gbm_proj_parallel(Y, M, subsample=subset, ncores=ncores, tol=tol,
                                 max_iter=max_iter, svd=svd, rng=rng, verbose=verbose)
    I, J = Y.shape
    LL = np.zeros(max_iter)
    loglik = []
    accepted = []
    lr_trace = []
    times = []
    t0 = _time.perf_counter()
    lr = 1.0                                                           # :84
    if batch is None:
        batch = np.ones(J, dtype=np.int64)
    batch = np.asarray(batch).astype(np.int64)
    nbatch = int(batch.max())                                          # :91
    bidx = batch - 1

    max_Y = Y.max()                                                    # :94
    nz = Y != 0                                                        # :95
    with np.errstate(divide="ignore"):
        log_csY = np.log(Y.sum(axis=0))                                # :96
        betas = log_csY.copy()                                         # :99
        log_rsy = np.zeros((nbatch, I))
        for k in range(nbatch):                                        # :103-105
            log_rsy[k, :] = np.log(Y[:, bidx == k].sum(axis=1))
        alphas = np.empty((I, nbatch))
        for k in range(nbatch):                                        # :106-108
            alphas[:, k] = log_rsy[k, :] - np.log(np.exp(betas[bidx == k]).sum())
    am = alphas.mean()
    betas = betas + am                                                 # :109
    alphas = alphas - am                                               # :110
    W = np.exp(alphas[:, bidx] + betas[None, :])                       # :111

    Z = (Y - W) / np.sqrt(W)                                           # :114
    u, d, v = _tsvd(Z, M, svd, rng)                                    # :115
    X = u @ (d[:, None] * v.T)                                         # :116
    X = np.sqrt(1.0 / W) * X                                           # :117
    X[X > 8] = 8                                                       # :119
    X[X < -8] = -8                                                     # :120
    Xt = np.zeros((I, J))                                              # :123
    LRA = None
    n_iter = 0
    hit_max = False
    alpha_xmax = None
    for i in range(1, max_iter + 1):                                   # :125
        n_iter = i
        alpha_xmax = np.abs(X).max(axis=1)   # diagnostic: max_j |X_ij| of the X this trip's alpha is solved from
        with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
            for k in range(nbatch):                                    # :127-130
                sel = bidx == k
                alphas[:, k] = log_rsy[k, :] - np.log(
                    np.exp(X[:, sel] + betas[sel][None, :]).sum(axis=1))
            am = alphas.mean()
            betas = betas + am                                         # :131
            alphas = alphas - am                                       # :132
            if infer_beta:                                             # :133-135
                betas = log_csY - np.log(np.exp(alphas[:, bidx] + X).sum(axis=0))
            W = np.exp(alphas[:, bidx] + X + betas[None, :])           # :136
            if time_by_iter:                                           # :139-142
                t1 = _time.perf_counter()
                times.append(t1 - t0)
                t0 = t1
            W[W > max_Y] = max_Y                                       # :145
            LL[i - 1] = (Y[nz] * np.log(W[nz])).sum() - W.sum()        # :151
            if _ll_inject and i in _ll_inject:
                LL[i - 1] = _ll_inject[i]
        lr_trace.append(lr)
        if math.isnan(LL[i - 1]) or math.isinf(LL[i - 1]):             # :152-157
            X = Xt
            lr = lr / 2
            accepted.append(False)
            continue
        if i >= 3:                                                     # :158
            tau = abs((LL[i - 1] - LL[i - 3]) / LL[i - 1])             # :159
            if tau < tol and lr <= 1.06 and i >= min_iter:             # :160
                accepted.append(False)
                break
            cmp_prev = LL[i - 2]
            if math.isnan(cmp_prev):
                raise RError("missing value where TRUE/FALSE needed")  # Q1
            if LL[i - 1] < cmp_prev:                                   # :164
                lr = max(lr / 2, 1.0)                                  # :165
                X = Xt                                                 # :166
                accepted.append(False)
                continue
            else:
                lr = lr * 1.05                                         # :169
        loglik.append(LL[i - 1])                                       # :174
        accepted.append(True)
        if verbose:
            print("Iteration: ", i, ". Objective=", "%.7g" % LL[i - 1])  # :175
        X, Xt, LRA = pgd_irlba(X, Xt, i, lr, W, Y, M, svd, rng)        # :178-180
        if i == max_iter:                                              # :182
            hit_max = True

    out = {}
    if return_W:                                                       # :189-191
        out["W"] = np.exp(alphas[:, bidx] + X) * np.exp(betas)[None, :]
    u, d, v = LRA                                                      # :192
    out["V"] = v
    out["D"] = d
    out["U"] = u
    out["loadings"] = u.copy()                                         # :196 (pre-flip, Q7)
    out["alpha"] = alphas.copy()
    out["beta"] = betas.copy()
    out["M"] = M
    if return_W:
        out["I"], out["J"] = out["W"].shape                            # :200 (Q9)
    out["obj"] = np.array(loglik)
    if time_by_iter:
        out["time"] = np.cumsum(times)                                 # :203
    out = process_results(out)                                         # :206
    # diagnostics (not part of the reference list)
    out["_LL"] = LL[:n_iter].copy()
    out["_accepted"] = np.array(accepted, dtype=bool)
    out["_lr"] = np.array(lr_trace)
    out["_n_iter"] = n_iter
    out["_hit_max_iter"] = hit_max
    out["_alpha_xmax"] = alpha_xmax
    return out


# ----------------------------------------------------------------------------
def poisson_glm_irls(Xd, y, offset, tol=1e-8, maxit=100):
    """fastglm(x, y, offset, family=poisson(), method=3) [ext]: IRLS with an LLT
    solve per step; start mu = y + 0.1 (poisson()$initialize); stop when
    |dev - devold| / (|dev| + 0.1) < tol; the standard errors come from the X'WX of
    the LAST solve (the weights that produced the returned coefficients), as in
    fastglm's save_se() and R's glm.fit/summary.glm.  Returns (coef, se, converged,
    n_solves)."""
    n, p = Xd.shape
    mu = y + 0.1
    eta = np.log(mu)

    def deviance(mu):
        with np.errstate(divide="ignore", invalid="ignore"):
            return 2.0 * np.sum(np.where(y > 0, y * np.log(y / mu), 0.0) - (y - mu))

    dev = deviance(mu)
    coef = np.zeros(p)
    A = None
    conv = False
    it = 0
    for it in range(1, maxit + 1):
        w = mu                                     # (dmu/deta)^2 / var = mu
        z = (eta - offset) + (y - mu) / mu
        XtW = Xd.T * w
        A = XtW @ Xd
        b = XtW @ z
        c, low = sla.cho_factor(A, lower=True, check_finite=False)
        coef = sla.cho_solve((c, low), b, check_finite=False)
        eta = Xd @ coef + offset
        mu = np.exp(eta)
        devold = dev
        dev = deviance(mu)
        if not np.isfinite(dev):
            raise FloatingPointError("non-finite deviance")
        if abs(dev - devold) / (abs(dev) + 0.1) < tol:
            conv = True
            break
    se = np.sqrt(np.diag(np.linalg.inv(A)))
    return coef, se, conv, it


def poisson_glm_newton(Xd, y, offset, tol=1e-8, maxit=30):
    """The iteration of the tensor-core projection path (scgbm_b200/csrc/proj_tc.cu), restated in float64 so that
    its ALGORITHM can be checked against fastglm's (poisson_glm_irls) without a GPU: Newton's method on the same
    likelihood (IRLS with the canonical link is Newton), started from the null model's MLE
    theta = (log(sum y / sum exp(offset)), 0, ..) instead of mu = y + 0.1, stopped on the Newton decrement
    g' H^-1 g / (|dev| + 0.1) < tol -- the predicted form of fastglm's |dev - devold| / (|dev| + 0.1) < tol -- with the
    step halved when the deviance rises; standard errors from the X'WX of the last solve, as fastglm reports them.
    Returns (coef, se, converged, n_solves)."""
    n, p = Xd.shape
    if y.sum() <= 0:
        raise FloatingPointError("no counts: handed to the exact iteration")
    theta = np.zeros(p)
    theta[0] = np.log(y.sum() / np.exp(offset).sum())
    theta_prev, dev_prev, solves, halv, se = theta.copy(), 0.0, 0, 0, np.zeros(p)
    ylogy = np.sum(np.where(y > 0, y * np.log(np.where(y > 0, y, 1.0)), 0.0))
    for _ in range(4 * maxit):
        eta = offset + Xd @ theta
        mu = np.exp(eta)
        dev = 2.0 * (ylogy - np.sum(y * eta) - np.sum(y - mu))
        if not np.isfinite(dev):
            raise FloatingPointError("non-finite deviance")
        if solves > 0 and dev > dev_prev + 1e-6 * (abs(dev_prev) + 0.1):
            halv += 1
            if halv > 12:
                raise FloatingPointError("step halving exhausted")
            theta = 0.5 * (theta + theta_prev)
            continue
        g = Xd.T @ (y - mu)
        H = (Xd.T * mu) @ Xd
        c, low = sla.cho_factor(H, lower=True, check_finite=False)
        delta = sla.cho_solve((c, low), g, check_finite=False)
        se = np.sqrt(np.diag(np.linalg.inv(H)))
        theta_prev, dev_prev, halv = theta.copy(), dev, 0
        theta = theta + delta
        solves += 1
        if float(g @ delta) / (abs(dev) + 0.1) < tol:
            return theta, se, True, solves
        if solves >= maxit:
            break
    return theta, se, False, solves


def gbm_proj_parallel(Y, M, subsample, ncores=1, tol=1e-4, max_iter=100, svd="exact",
                      rng=None, verbose=False, glm_tol=1e-8, _subfit=None):
    """R/fit.R:213-285.  `subsample` scalar -> `rng.choice` (R's sample() stream is
    not restated); vector -> 1-based column indices as in R (:221-223)."""
    Y = np.asarray(Y, dtype=np.float64)
    I, J = Y.shape
    with np.errstate(divide="ignore"):
        alphas_full = np.log(Y.sum(axis=1))                            # :217
    if np.ndim(subsample) == 0:                                        # :218-220
        rng = rng or np.random.default_rng(0)
        jxs = rng.choice(J, size=int(subsample), replace=False)
    else:
        jxs = np.asarray(subsample, dtype=np.int64) - 1
    Y_sub = Y[:, jxs]
    ixs = np.nonzero(Y_sub.sum(axis=1) > 5)[0]                         # :225
    Y_sub = Y_sub[ixs, :]
    if _subfit is not None:      # test hook: reuse a given subsample fit (isolates the projection step)
        out = dict(_subfit)
    else:
        out = gbm_sc(Y_sub, M=M, tol=tol, max_iter=max_iter, svd=svd, rng=rng, verbose=verbose)  # :227
    U = np.hstack([np.ones((len(ixs), 1)), out["U"]])                  # :229-232
    alpha = out["alpha"][:, 0]                                         # :234
    Yf = Y[ixs, :]                                                     # :237
    V = np.zeros((J, 2 * M + 1))                                       # :238
    fail = np.zeros(J, dtype=bool)
    for j in range(J):                                                 # :248-264
        try:
            coef, se, _, _ = poisson_glm_irls(U, Yf[:, j], alpha, tol=glm_tol)
            V[j, : M + 1] = coef
            V[j, M + 1:] = se[1:]
        except Exception:                                              # try({...}) :251
            fail[j] = True
    out["se_scores"] = V[:, M + 1: 2 * M + 1]                          # :268
    out["scores"] = V[:, 1: M + 1]                                     # :269
    beta = V[:, 0].copy()                                              # :270
    alpha_f = np.zeros(I)                                              # :272
    alpha_f[ixs] = out["alpha"][:, 0]                                  # :273
    mask = np.ones(I, dtype=bool)
    mask[ixs] = False
    with np.errstate(over="ignore"):
        alpha_f[mask] = alphas_full[mask] - np.log(np.exp(beta).sum())  # :274
    am = alpha_f.mean()
    out["beta"] = beta + am                                            # :275
    out["alpha"] = alpha_f - am                                        # :276-277
    Uf = np.zeros((I, M))                                              # :278-280
    Uf[ixs, :] = out["U"]
    out["U"] = Uf
    out["V"] = None                                                    # :281
    out["loadings"] = None   # :282 assigns stats::loadings (a function) -- a reference bug
    out["_ixs"] = ixs
    out["_jxs"] = jxs
    out["_fail"] = fail
    return out


# ----------------------------------------------------------------------------
def get_se(out, EPS=0.0):
    """R/uncertainty.R:14-35."""
    M = len(out["D"])
    U = out["U"]
    V = out["scores"]
    W = out["W"]
    I, J = W.shape
    eye = EPS * np.eye(M)
    se_V = np.empty((J, M))
    for j in range(J):                                                 # :22-25
        Mat = U.T @ (W[:, j][:, None] * U)
        se_V[j, :] = np.sqrt(np.diag(np.linalg.inv(Mat + eye)))
    se_U = np.empty((I, M))
    for i in range(I):                                                 # :29-32
        Mat = V.T @ (W[i, :][:, None] * V)
        se_U[i, :] = np.sqrt(np.diag(np.linalg.inv(Mat + eye)))
    out = dict(out)
    out["se_scores"] = se_V                                            # :26
    out["se_loadings"] = se_U                                          # :33
    return out


def get_se_fast(out, EPS=0.0):
    """Same numbers as get_se, vectorised (for C2-size parity runs)."""
    M = len(out["D"])
    U, V, W = out["U"], out["scores"], out["W"]
    eye = EPS * np.eye(M)
    KU = np.einsum("ia,ib->iab", U, U).reshape(U.shape[0], M * M)
    G = (W.T @ KU).reshape(-1, M, M) + eye
    se_V = np.sqrt(np.diagonal(np.linalg.inv(G), axis1=1, axis2=2))
    KV = np.einsum("ja,jb->jab", V, V).reshape(V.shape[0], M * M)
    G = (W @ KV).reshape(-1, M, M) + eye
    se_U = np.sqrt(np.diagonal(np.linalg.inv(G), axis1=1, axis2=2))
    out = dict(out)
    out["se_scores"] = se_V
    out["se_loadings"] = se_U
    return out


# ----------------------------------------------------------------------------
# synthetic inputs following R/simulate.R (harness RNG, not R's stream)
# ----------------------------------------------------------------------------
def sim_data(I, J, d, alpha_mean=0.0, beta_mean=0.0, K=2.0, seed=1, nb_theta=None):
    """simData recipe, R/simulate.R:24-48 (Haar U,V via QR of a Gaussian);
    nb_theta -> negative-binomial counts as in simNullNB (R/simulate.R:60-65)."""
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((I, d)))
    V, _ = np.linalg.qr(rng.standard_normal((J, d)))
    D = np.linspace(K * (math.sqrt(I) + math.sqrt(J)), math.sqrt(I) + math.sqrt(J), d)
    for m in range(d):
        if U[0, m] < 0:
            U[:, m] = -U[:, m]
            V[:, m] = -V[:, m]
    alpha = rng.normal(alpha_mean, 1.0, I)
    beta = rng.normal(beta_mean, 1.0, J)
    Lam = alpha[:, None] + beta[None, :] + (U * D) @ V.T
    mu = np.exp(Lam)
    if nb_theta is None:
        Y = rng.poisson(mu)
    else:
        g = rng.gamma(nb_theta, mu / nb_theta)
        Y = rng.poisson(g)
    keep_i = Y.sum(axis=1) > 0      # Q11 precondition
    keep_j = Y.sum(axis=0) > 0
    Y = Y[keep_i][:, keep_j]
    return dict(Y=np.asfortranarray(Y.astype(np.int32)), U=U[keep_i], V=(V * D)[keep_j], D=D,
                alpha=alpha[keep_i], beta=beta[keep_j])
