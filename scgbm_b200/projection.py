"""gbm.proj.parallel (R/fit.R:213-285): fit on a subsample of cells, then project
every cell onto the fitted loadings with one Poisson GLM per cell.  The subsample
draw, row filter and re-centring stay on the host like the reference's R code; the
fit and the J per-cell GLMs run on the GPU (scgbm_fit, scgbm_project)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def gbm_proj_parallel(Y, M, subsample=2000, min_counts=5, ncores=1, tol=1e-4, max_iter=100, *,
                      device=-1, rng=None, profile=False, svd_tol=0.0, seed=0, glm_tol=1e-8, glm_maxit=100,
                      glm_method="auto", comm=None, subsample_Y=None):
    """Cells sharded over ranks (BASELINE configs[3]): pass `comm` and the
    subsample's count matrix `subsample_Y` (I x s, the same on every rank -- the host
    harness gathers those columns).  Y then holds this rank's cells only; every rank
    runs the identical subsample fit (bit-reproducible), projects its own cells, and the
    only exchanges are the gene row sums and the scalar sum(exp(beta)) of R/fit.R:274.

    glm_method: "exact" (one fp64 IRLS per cell, fastglm's iteration), "tcgen05" (all cells in lock step on the tensor
    cores, include/scgbm.h scgbm_glm_method) or "auto" (tcgen05 for integer counts and >= 4096 cells)."""
    from .api import gbm_sc, _as_counts, _is_sparse, _NP2DT, _canonical_csc
    lib = L.load()
    if (comm is None) != (subsample_Y is None):
        raise ValueError("comm and subsample_Y go together")
    sparse = _is_sparse(Y)
    if sparse:                       # dgCMatrix-style input stays sparse on the host (R/fit.R:224 densifies only Y.sub)
        Ys = _canonical_csc(Y)
        I, J = Ys.shape
        rs = np.asarray(Ys.sum(axis=1), dtype=np.float64).ravel()
    else:
        Yh, y_dt = _as_counts(Y)
        I, J = Yh.shape
        rs = None      # rowSums(Y) (R/fit.R:217) comes back from scgbm_project, computed from the same upload:
                       # sweeping a 20k x 1M host matrix with numpy would take longer than the whole projection
    if subsample_Y is not None:
        jxs = None
    elif np.ndim(subsample) == 0:                                            # R/fit.R:218-220
        rng = rng or np.random.default_rng()
        jxs = rng.choice(J, size=int(subsample), replace=False)            # sample(1:J, size)
    else:
        jxs = np.asarray(subsample, dtype=np.int64) - 1                    # 1-based like R (:221-223)
    if subsample_Y is not None:
        Y_sub = subsample_Y.toarray() if _is_sparse(subsample_Y) else np.asarray(subsample_Y)
        if Y_sub.shape[0] != I:
            raise ValueError("subsample_Y must have the same genes as Y")
    else:
        Y_sub = np.asfortranarray(Ys[:, jxs].toarray() if sparse else Yh[:, jxs])   # as.matrix(Y.sub)  :224
    ixs = np.nonzero(Y_sub.sum(axis=1, dtype=np.float64) > 5)[0]           # hard-coded 5  :225
    if len(ixs) < Y_sub.shape[0]:                                          # Y.sub[ixs, ]  :226
        # rows of a column-major matrix: gather columns of its (row-major) transposed view -- a plain
        # `Y_sub[ixs, :]` goes through numpy's generic fancy-indexing path, 10x slower at 20k x 50k
        Y_sub = np.take(np.asfortranarray(Y_sub).T, ixs, axis=1).T
    Y_sub = np.asfortranarray(Y_sub)
    out = gbm_sc(Y_sub, M, tol=tol, max_iter=max_iter, device=device, profile=profile,
                 diagnostics=profile, svd_tol=svd_tol, seed=seed)          # :227
    Uk = np.asfortranarray(out["U"])                                       # :229
    alpha_k = np.ascontiguousarray(out["alpha"][:, 0])                     # :234
    ixs64 = np.ascontiguousarray(ixs.astype(np.int64))
    beta = np.zeros(J); scores = np.zeros((J, M), order="F"); se_scores = np.zeros((J, M), order="F")
    fail = np.zeros(J, dtype=np.int8); iters = np.zeros(J, dtype=np.int32)
    a = L.ProjArgs()
    if sparse:
        xs = np.ascontiguousarray(Ys.data)
        if xs.dtype not in _NP2DT:
            xs = xs.astype(np.float64)
        ci = np.ascontiguousarray(Ys.indices, dtype=np.int32); cp = np.ascontiguousarray(Ys.indptr, dtype=np.int64)
        a.Y = None; a.csc_i = L.ptr(ci); a.csc_p = L.ptr(cp); a.csc_x = L.ptr(xs); a.csc_x_dtype = _NP2DT[xs.dtype]
        a.y_dtype = 0; a.y_on_device = 0; a.I_full = I; a.J = J; a.ldY = I
    else:
        a.Y = L.ptr(Yh); a.y_dtype = y_dt; a.y_on_device = 0; a.I_full = I; a.J = J; a.ldY = I
    a.ixs = L.ptr(ixs64); a.Ik = len(ixs64); a.M = int(M); a.U = L.ptr(Uk); a.alpha = L.ptr(alpha_k)
    a.glm_tol = float(glm_tol); a.glm_maxit = int(glm_maxit); a.device = int(device); a.profile = int(bool(profile))
    a.glm_method = {"auto": 0, "exact": 1, "tcgen05": 2}[glm_method]
    o = L.ProjOut()
    o.beta = L.ptr(beta); o.scores = L.ptr(scores); o.se_scores = L.ptr(se_scores)
    o.fail = L.ptr(fail); o.iters = L.ptr(iters)
    rs_dev = np.zeros(I) if rs is None else None
    o.rowsum_full = L.ptr(rs_dev)
    L.check(lib.scgbm_project(C.byref(a), C.byref(o)))
    if rs is None:
        rs = rs_dev
    if comm is not None:
        rs = comm.allreduce(rs)
    with np.errstate(divide="ignore"):
        alphas_full = np.log(rs)                                           # R/fit.R:217
    out["se_scores"] = se_scores                                           # :268
    out["scores"] = scores                                                 # :269
    alpha = np.zeros(I)                                                    # :272
    alpha[ixs] = out["alpha"][:, 0]                                        # :273
    mask = np.ones(I, dtype=bool); mask[ixs] = False
    with np.errstate(over="ignore"):
        sum_eb = np.exp(beta).sum()
        if comm is not None:
            sum_eb = float(comm.allreduce(np.array([sum_eb]))[0])
        alpha[mask] = alphas_full[mask] - np.log(sum_eb)                   # :274
    am = alpha.mean()
    out["beta"] = beta + am                                                # :275
    out["alpha"] = alpha - am                                              # :276-277
    Uf = np.zeros((I, M), order="F"); Uf[ixs, :] = out["U"]                # :278-280
    out["U"] = Uf
    out["V"] = None                                                        # :281
    out["loadings"] = None   # :282 assigns the undefined `loadings` (stats::loadings): a reference bug
    out["_ixs"] = ixs; out["_jxs"] = jxs; out["_fail"] = fail.astype(bool); out["_glm_iters"] = iters
    out["_glm_method"] = {2: "tcgen05"}.get(o.method_used, "exact")
    out["_glm_tc"] = dict(sweeps=o.tc_sweeps, stragglers=o.tc_stragglers)
    if profile:
        out["_prof_proj"] = o.prof.as_dict()
    return out
