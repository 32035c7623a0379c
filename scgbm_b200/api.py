"""Host-side mirror of the reference's R interface for the hot path
(`gbm.sc`, R/fit.R:52-62; `get.se`, R/uncertainty.R:14) on top of the C ABI.

R is not installed in this image, so this module plays the role the R wrapper in
`r-package/` plays for R users: same argument names and meaning, same returned
list (as an ordered dict with R's element order, quirk Q13), same error behaviour
(`RError` for `stop()`, `warnings.warn` for `warning()`).  All numerics happen in
libscgbm.so on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import sys
import warnings
from collections import OrderedDict

import numpy as np

from . import _lib as L


class RError(L.ScgbmError):
    """An R `stop()` raised by the reference for the same input."""


_NP2DT = {np.dtype(np.int32): L.I32, np.dtype(np.float64): L.F64, np.dtype(np.uint8): L.U8,
          np.dtype(np.uint16): L.U16, np.dtype(np.float32): L.F32}


def _as_counts(Y):
    """Column-major view acceptable to the C ABI (R matrices are column-major)."""
    Y = np.asarray(Y)
    if Y.ndim != 2:
        raise RError(1, "Y must be a matrix")
    if Y.dtype not in _NP2DT:
        if np.issubdtype(Y.dtype, np.integer) or Y.dtype == np.bool_:
            Y = Y.astype(np.int32)
        else:
            Y = Y.astype(np.float64)
    if not Y.flags.f_contiguous:
        Y = np.asfortranarray(Y)
    return Y, _NP2DT[Y.dtype]


def _is_sparse(Y):
    try:
        import scipy.sparse as sp
        return sp.issparse(Y)
    except Exception:       # pragma: no cover
        return False


def _batch_codes(batch, J, comm=None, levels=None):
    """R: `batch` must be a factor (R/fit.R:86-88); here: any 1-D array of labels.
    Returns 1-based int32 codes = as.integer(factor(x, levels = sorted unique labels)) and the level count.

    With `comm` (cells sharded over ranks) the levels are the union over ALL ranks: a shard that lacks a level
    must still give its cells the same code as everybody else (include/scgbm.h: codes are globally consistent),
    otherwise row sums, log.rsy and alpha[, batch] silently mix batches.  The union is agreed over the
    library's own transport (Comm.allgather_bytes); explicit `levels` skip the exchange."""
    if batch is None:
        if comm is not None:      # all ranks must agree that there is no batch argument
            flags = comm.allgather_bytes(b"0")
            if any(f != b"0" for f in flags):
                raise RError(1, "batch must be given on every rank or on none")
        return None, 1
    b = np.asarray(batch)
    if b.shape != (J,):
        raise RError(1, "Batch must be encoded as factor.")
    if levels is None:
        levels = np.unique(b)
        if comm is not None:
            import pickle
            parts = comm.allgather_bytes(b"1" + pickle.dumps(levels))
            if any(p[:1] != b"1" for p in parts):
                raise RError(1, "batch must be given on every rank or on none")
            try:
                levels = np.unique(np.concatenate([pickle.loads(p[1:]) for p in parts]))
            except Exception as e:       # labels of different types on different ranks
                raise RError(1, f"batch levels cannot be agreed across ranks: {e}")
    else:
        levels = np.unique(np.asarray(levels))
    codes = np.searchsorted(levels, b)
    ok = (codes < len(levels))
    ok[ok] &= levels[codes[ok]] == b[ok]
    if not ok.all():
        raise RError(1, "batch contains labels that are not among the levels")
    return np.ascontiguousarray(codes.astype(np.int32) + 1), int(len(levels))


def _canonical_csc(Y):
    """dgCMatrix semantics for scipy input: duplicate (i, j) entries are SUMMED (Matrix and scipy both do) and
    row indices sorted.  The scatter kernel writes entries, it does not add them, so duplicates must not reach
    it (a last-writer-wins race would silently lose counts)."""
    Ys = Y.tocsc()
    if not Ys.has_canonical_format:
        Ys = Ys.copy()
        Ys.sum_duplicates()
    return Ys


def gbm_sc(Y, M, max_iter=100, tol=1e-4, subset=None, ncores=1, infer_beta=False, return_W=True,
           batch=None, time_by_iter=False, min_iter=30, *, device=-1, comm=None, y_storage=L.STORE_AUTO,
           svd_extra=0, svd_depth=0, svd_tol=0.0, seed=0, profile=False, verbose=False, quiet=True,
           diagnostics=False, y_device_ptr=None, y_device_shape=None, y_device_dtype=None, rng=None,
           keep_factors=False, subset_Y=None, ngpu=1, devices=None, batch_levels=None, glm_method="auto"):
    """gbm.sc(Y, M, max.iter, tol, subset, ncores, infer.beta, return.W, batch,
    time.by.iter, min.iter)   -- R/fit.R:52-62.

    Extensions (keyword only): `device`, `comm` (one NCCL rank per GPU; Y then holds
    this rank's cells), `y_storage`, `svd_*`, `seed`, `profile`, `diagnostics`
    (adds `_LL`, `_accepted`, `_n_iter`, `_prof`), `y_device_ptr` (Y already
    resident in HBM in the library's own layout), `keep_factors` (adds `_Xu`, `_Xv`,
    `_batch`: the factors of the X that alpha/beta/W belong to, so that `get_se` works
    after `return_W=False` -- the reference cannot, quirk Q9), `ngpu` (>= 2: ONE call uses that many GPUs of
    the box -- the library splits the cells over one host thread / NCCL rank per GPU; all outputs cover all
    cells), `batch_levels` (explicit factor levels; with `comm` they are otherwise agreed across ranks),
    `glm_method` (projection only: "auto" | "exact" | "tcgen05", see gbm_proj_parallel).
    """
    if subset is not None or subset_Y is not None:             # R/fit.R:67-73
        from .projection import gbm_proj_parallel
        out = gbm_proj_parallel(Y, M, subsample=subset if subset is not None else 0, ncores=ncores, tol=tol, max_iter=max_iter,
                                device=device, rng=rng, profile=profile, svd_tol=svd_tol, seed=seed,
                                comm=comm, subsample_Y=subset_Y, glm_method=glm_method)
        if not quiet:
            print("For users of newer versions (1.0.1+): the `scores` matrix now contains factor scores.",
                  file=sys.stderr)
        return out
    lib = L.load()
    csc = None
    if _is_sparse(Y):
        # Matrix::dgCMatrix-style input: handed over column-compressed, densified on the GPU
        Ys = _canonical_csc(Y)
        I, J = Ys.shape
        xs = np.ascontiguousarray(Ys.data)
        if xs.dtype not in _NP2DT:
            xs = xs.astype(np.float64)
        csc = (np.ascontiguousarray(Ys.indices, dtype=np.int32), np.ascontiguousarray(Ys.indptr, dtype=np.int64), xs)
        y_ptr, y_dt, on_dev, ldY, Yh = None, _NP2DT[xs.dtype], 0, I, None
    elif y_device_ptr is not None:
        I, J = y_device_shape
        y_ptr, y_dt, on_dev, ldY = C.c_void_p(y_device_ptr), y_device_dtype, 1, ((I + 63) // 64) * 64
        Yh = None
    else:
        Yh, y_dt = _as_counts(Y)
        I, J = Yh.shape
        y_ptr, on_dev, ldY = L.ptr(Yh), 0, I
    if M < 1:
        raise RError(1, "M must be >= 1")                      # R/fit.R:296-298
    M = int(M)
    codes, nbatch = _batch_codes(batch, J, comm=comm, levels=batch_levels)
    ngpu = int(ngpu or 1)
    if ngpu > 1 and (comm is not None or y_device_ptr is not None):
        raise RError(1, "ngpu > 1 needs Y in host memory and no comm (comm = one process per GPU)")
    dev_list = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
    if dev_list is not None and len(dev_list) != ngpu:
        raise RError(1, "devices must list ngpu CUDA ordinals")
    max_iter = int(max_iter)
    U = np.zeros((I, M), order="F"); V = np.zeros((J, M), order="F"); D = np.zeros(M)
    loadings = np.zeros((I, M), order="F"); scores = np.zeros((J, M), order="F")
    alpha = np.zeros((I, nbatch), order="F"); beta = np.zeros(J)
    obj = np.zeros(max_iter); LLv = np.zeros(max_iter); acc = np.zeros(max_iter, dtype=np.int8)
    tvec = np.zeros(max_iter) if time_by_iter else None
    W = np.zeros((I, J), order="F") if return_W else None

    def _cb(it, objective, _user):
        if not quiet:
            print("Iteration: ", it, ". Objective=", "%.7g" % objective)   # R/fit.R:175
    cb = L.PROGRESS_CB(_cb)

    a = L.FitArgs()
    a.Y = y_ptr; a.y_dtype = y_dt; a.y_on_device = on_dev; a.I = I; a.J = J; a.ldY = ldY
    a.M = M; a.max_iter = max_iter; a.tol = float(tol); a.min_iter = int(min_iter)
    a.infer_beta = int(bool(infer_beta)); a.batch = L.ptr(codes); a.nbatch = nbatch
    a.return_W = int(bool(return_W)); a.time_by_iter = int(bool(time_by_iter))
    a.y_storage = int(y_storage); a.device = int(device)
    a.comm = comm.handle if comm is not None else None
    a.svd_extra = int(svd_extra); a.svd_depth = int(svd_depth); a.svd_tol = float(svd_tol)
    a.seed = int(seed); a.profile = int(profile); a.verbose = int(bool(verbose))   # profile: 0 / 1 all classes / 2 fused pass only
    a.progress = cb; a.progress_user = None
    if csc is not None:
        a.csc_i = L.ptr(csc[0]); a.csc_p = L.ptr(csc[1]); a.csc_x = L.ptr(csc[2]); a.csc_x_dtype = y_dt
    a.ngpu = ngpu; a.devices = L.ptr(dev_list)
    o = L.FitOut()
    o.U = L.ptr(U); o.D = L.ptr(D); o.V = L.ptr(V); o.loadings = L.ptr(loadings); o.scores = L.ptr(scores)
    o.alpha = L.ptr(alpha); o.beta = L.ptr(beta); o.obj = L.ptr(obj); o.time = L.ptr(tvec)
    o.W = L.ptr(W); o.LL = L.ptr(LLv); o.accepted = L.ptr(acc)
    Xu = np.zeros((I, M), order="F") if keep_factors else None
    Xv = np.zeros((J, M), order="F") if keep_factors else None
    o.Xu = L.ptr(Xu); o.Xv = L.ptr(Xv)
    st = lib.scgbm_fit(C.byref(a), C.byref(o))
    if st == 1:
        raise RError(st, (lib.scgbm_last_error() or b"").decode())
    L.check(st)
    if o.hit_max_iter:                                          # R/fit.R:182-185
        warnings.warn("Maximum number of iterations reached (increase max.iter).\n"
                      "              Possible non-convergence.")
    if o.svd_unconverged:                                       # irlba would warn "did not converge" [ext]
        warnings.warn(f"truncated SVD: {o.svd_unconverged} call(s) returned with a residual estimate above svd_tol "
                      f"(worst {o.svd_worst_est:.2e}); the projection of those iterations is approximate.")
    out = OrderedDict()                                         # element order: quirk Q13
    if return_W:
        out["W"] = W
    out["V"] = V; out["D"] = D; out["U"] = U; out["loadings"] = loadings
    out["alpha"] = alpha; out["beta"] = beta; out["M"] = M
    if return_W:                                                # Q9: I, J only exist with W
        out["I"] = I; out["J"] = J
    out["obj"] = obj[: o.n_obj].copy()
    if time_by_iter:
        out["time"] = tvec[: o.n_iter].copy()
    out["scores"] = scores
    if not quiet:
        print("For users of newer versions (1.0.1+): the `scores` matrix now contains factor scores, "
              "the `V` matrix is UNSCALED scores.", file=sys.stderr)   # R/fit.R:209
    if keep_factors:
        out["_x_rank"] = int(o.x_rank)
        out["_Xu"] = Xu; out["_Xv"] = Xv; out["_batch"] = codes
    if diagnostics or profile:
        out["_LL"] = LLv[: o.n_iter].copy()
        out["_accepted"] = acc[: o.n_iter].astype(bool)
        out["_n_iter"] = int(o.n_iter)
        out["_hit_max_iter"] = bool(o.hit_max_iter)
        out["_y_storage"] = int(o.y_storage)
        out["_max_Y"] = float(o.max_Y)
        out["_prof"] = o.prof.as_dict()
        out["_svd_unconverged"] = int(o.svd_unconverged)
        out["_svd_worst_est"] = float(o.svd_worst_est)
        out["_underflow_reruns"] = int(o.underflow_reruns)
        out["_exact_rows"] = int(o.exact_rows)
    return out


def get_se(out, EPS=0.0, *, device=-1, comm=None, profile=False, ngpu=1, devices=None):
    """get.se(out, EPS)  -- R/uncertainty.R:14-35.  Needs out['W'] like the reference
    (quirk Q9), or -- extension -- the `keep_factors=True` outputs of gbm_sc, from which W is
    re-evaluated on the device chunk by chunk (the only option at sizes where W cannot exist)."""
    lib = L.load()
    U = np.asfortranarray(out["U"], dtype=np.float64)
    S = np.asfortranarray(out["scores"], dtype=np.float64)
    M = len(out["D"])
    a = L.SeArgs()
    keep = []
    if out.get("W") is not None:
        W = np.asfortranarray(out["W"], dtype=np.float64)
        I, J = W.shape
        a.W = L.ptr(W); a.nbatch = 1
    elif out.get("_Xu") is not None and out.get("_x_rank", -1) >= 0:
        I, J = U.shape[0], S.shape[0]
        al = np.asfortranarray(out["alpha"], dtype=np.float64).reshape(I, -1, order="F")
        be = np.ascontiguousarray(out["beta"], dtype=np.float64)
        Xu = np.asfortranarray(out["_Xu"], dtype=np.float64); Xv = np.asfortranarray(out["_Xv"], dtype=np.float64)
        bt = None if out.get("_batch") is None else np.ascontiguousarray(out["_batch"], dtype=np.int32)   # 1-based, R/fit.R:90
        keep = [al, be, Xu, Xv, bt]
        a.W = None; a.alpha = L.ptr(al); a.beta = L.ptr(be); a.batch = L.ptr(bt); a.nbatch = al.shape[1]
        a.Xu = L.ptr(Xu); a.Xv = L.ptr(Xv); a.Mx = int(out["_x_rank"])
    else:
        raise RError(1, "get.se needs out$W: fit with return.W = TRUE (R/uncertainty.R:18) or keep_factors=True")
    se_s = np.zeros((J, M), order="F"); se_l = np.zeros((I, M), order="F")
    a.I = I; a.J = J; a.M = M; a.U = L.ptr(U); a.scores = L.ptr(S)
    a.comm = comm.handle if comm is not None else None
    a.EPS = float(EPS); a.device = int(device); a.profile = int(bool(profile))
    dev_list = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
    a.ngpu = int(ngpu or 1); a.devices = L.ptr(dev_list)
    o = L.SeOut()
    o.se_scores = L.ptr(se_s); o.se_loadings = L.ptr(se_l)
    st = lib.scgbm_get_se(C.byref(a), C.byref(o))
    if st in (1, 3):
        raise RError(st, (lib.scgbm_last_error() or b"").decode())
    L.check(st)
    res = OrderedDict(out)
    res["se_scores"] = se_s                                     # R/uncertainty.R:26
    res["se_loadings"] = se_l                                   # R/uncertainty.R:33
    if profile:
        res["_prof_se"] = o.prof.as_dict()
    return res
