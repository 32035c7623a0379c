"""ctypes bindings of libscgbm.so (include/scgbm.h).  No torch, no fallback: if the
CUDA library is missing or no device is present every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscgbm.so")

I32, F64, U8, U16, F32 = 0, 1, 2, 3, 4
STORE_AUTO, STORE_U8, STORE_U16, STORE_F32 = 0, 1, 2, 3
K_NCLASS = 11
K_NAMES = ["pack", "rowsum", "colsum", "fused", "prod_n", "prod_t", "small", "se", "proj", "comm", "misc"]
UNIQUE_ID_BYTES = 128


class ScgbmError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * K_NCLASS), ("launches", C.c_int64 * K_NCLASS),
                ("fused_bytes", C.c_double), ("total_ms", C.c_double),
                ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
                ("svd_passes", C.c_int64), ("svd_calls", C.c_int64)]

    def as_dict(self):
        return {
            "ms": {K_NAMES[k]: self.ms[k] for k in range(K_NCLASS)},
            "launches": {K_NAMES[k]: int(self.launches[k]) for k in range(K_NCLASS)},
            "fused_bytes": self.fused_bytes, "total_ms": self.total_ms,
            "h2d_bytes": self.h2d_bytes, "d2h_bytes": self.d2h_bytes,
            "svd_passes": int(self.svd_passes), "svd_calls": int(self.svd_calls),
        }


PROGRESS_CB = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_void_p)
DP = C.POINTER(C.c_double)


class FitArgs(C.Structure):
    _fields_ = [("Y", C.c_void_p), ("y_dtype", C.c_int32), ("y_on_device", C.c_int32),
                ("I", C.c_int64), ("J", C.c_int64), ("ldY", C.c_int64),
                ("M", C.c_int32), ("max_iter", C.c_int32), ("tol", C.c_double),
                ("min_iter", C.c_int32), ("infer_beta", C.c_int32),
                ("batch", C.c_void_p), ("nbatch", C.c_int32), ("return_W", C.c_int32),
                ("time_by_iter", C.c_int32),
                ("y_storage", C.c_int32), ("device", C.c_int32), ("comm", C.c_void_p),
                ("svd_extra", C.c_int32), ("svd_depth", C.c_int32), ("svd_tol", C.c_double),
                ("seed", C.c_uint64), ("profile", C.c_int32), ("verbose", C.c_int32),
                ("progress", PROGRESS_CB), ("progress_user", C.c_void_p),
                ("csc_i", C.c_void_p), ("csc_p", C.c_void_p), ("csc_x", C.c_void_p), ("csc_x_dtype", C.c_int32),
                ("ngpu", C.c_int32), ("devices", C.c_void_p)]


class FitOut(C.Structure):
    _fields_ = [("U", C.c_void_p), ("D", C.c_void_p), ("V", C.c_void_p), ("loadings", C.c_void_p),
                ("scores", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p),
                ("obj", C.c_void_p), ("time", C.c_void_p), ("W", C.c_void_p),
                ("LL", C.c_void_p), ("accepted", C.c_void_p),
                ("n_obj", C.c_int32), ("n_iter", C.c_int32), ("hit_max_iter", C.c_int32),
                ("y_storage", C.c_int32), ("max_Y", C.c_double), ("prof", Profile),
                ("Xu", C.c_void_p), ("Xv", C.c_void_p), ("x_rank", C.c_int32),
                ("svd_unconverged", C.c_int32), ("underflow_reruns", C.c_int32), ("svd_worst_est", C.c_double),
                ("exact_rows", C.c_int64)]


class SeArgs(C.Structure):
    _fields_ = [("I", C.c_int64), ("J", C.c_int64), ("M", C.c_int32),
                ("U", C.c_void_p), ("scores", C.c_void_p), ("W", C.c_void_p),
                ("alpha", C.c_void_p), ("beta", C.c_void_p), ("batch", C.c_void_p),
                ("nbatch", C.c_int32), ("Xu", C.c_void_p), ("Xv", C.c_void_p), ("Mx", C.c_int32),
                ("EPS", C.c_double), ("device", C.c_int32), ("comm", C.c_void_p), ("profile", C.c_int32),
                ("ngpu", C.c_int32), ("devices", C.c_void_p)]


class SeOut(C.Structure):
    _fields_ = [("se_scores", C.c_void_p), ("se_loadings", C.c_void_p), ("n_singular", C.c_int64),
                ("prof", Profile)]


class ProjArgs(C.Structure):
    _fields_ = [("Y", C.c_void_p), ("y_dtype", C.c_int32), ("y_on_device", C.c_int32),
                ("I_full", C.c_int64), ("J", C.c_int64), ("ldY", C.c_int64),
                ("ixs", C.c_void_p), ("Ik", C.c_int64), ("M", C.c_int32),
                ("U", C.c_void_p), ("alpha", C.c_void_p), ("glm_tol", C.c_double),
                ("glm_maxit", C.c_int32), ("device", C.c_int32), ("profile", C.c_int32),
                ("csc_i", C.c_void_p), ("csc_p", C.c_void_p), ("csc_x", C.c_void_p), ("csc_x_dtype", C.c_int32),
                ("glm_method", C.c_int32)]


class ProjOut(C.Structure):
    _fields_ = [("beta", C.c_void_p), ("scores", C.c_void_p), ("se_scores", C.c_void_p),
                ("fail", C.c_void_p), ("iters", C.c_void_p), ("prof", Profile), ("rowsum_full", C.c_void_p),
                ("method_used", C.c_int32), ("tc_sweeps", C.c_int32), ("tc_stragglers", C.c_int64)]


class SimArgs(C.Structure):
    _fields_ = [("I", C.c_int64), ("J", C.c_int64), ("ld", C.c_int64), ("j_offset", C.c_int64),
                ("d", C.c_int32), ("U", C.c_void_p), ("V", C.c_void_p), ("D", C.c_void_p),
                ("alpha", C.c_void_p), ("beta", C.c_void_p), ("nb_theta", C.c_double),
                ("seed", C.c_uint64), ("out_dtype", C.c_int32), ("device", C.c_int32)]


class SimDataArgs(C.Structure):
    _fields_ = [("model", C.c_int32), ("d", C.c_int32), ("I", C.c_int64), ("J", C.c_int64), ("J_total", C.c_int64),
                ("j_offset", C.c_int64), ("alpha_mean", C.c_double), ("beta_mean", C.c_double), ("K", C.c_double),
                ("theta", C.c_double), ("seed", C.c_uint64), ("out_dtype", C.c_int32), ("y_on_device", C.c_int32),
                ("ldY", C.c_int64), ("device", C.c_int32)]


class SimDataOut(C.Structure):
    _fields_ = [("Y", C.c_void_p), ("U", C.c_void_p), ("V", C.c_void_p), ("D", C.c_void_p), ("alpha", C.c_void_p),
                ("beta", C.c_void_p), ("n_saturated", C.c_int64)]


class CciArgs(C.Structure):
    _fields_ = [("J", C.c_int64), ("M", C.c_int32), ("scores", C.c_void_p), ("se_scores", C.c_void_p),
                ("reps", C.c_int32), ("rep_offset", C.c_int32), ("seed", C.c_uint64), ("out_on_device", C.c_int32),
                ("device", C.c_int32)]


class TsvdArgs(C.Structure):
    _fields_ = [("I", C.c_int64), ("J", C.c_int64), ("R", C.c_void_p), ("c", C.c_double),
                ("nterms", C.c_int32), ("A", C.c_void_p * 2), ("d", C.c_void_p * 2), ("B", C.c_void_p * 2),
                ("Mt", C.c_int32 * 2), ("coef", C.c_double * 2), ("M", C.c_int32),
                ("svd_extra", C.c_int32), ("svd_depth", C.c_int32), ("svd_tol", C.c_double),
                ("V0", C.c_void_p), ("seed", C.c_uint64), ("device", C.c_int32),
                ("kernel_variant", C.c_int32)]


class TsvdOut(C.Structure):
    _fields_ = [("U", C.c_void_p), ("D", C.c_void_p), ("V", C.c_void_p), ("passes", C.c_int32),
                ("est", C.c_double)]


class ProdArgs(C.Structure):
    _fields_ = [("I", C.c_int64), ("J", C.c_int64), ("R", C.c_void_p), ("P", C.c_void_p), ("Q", C.c_void_p),
                ("b", C.c_int32), ("c", C.c_double), ("variant", C.c_int32), ("reps", C.c_int32),
                ("device", C.c_int32)]


class ProdOut(C.Structure):
    _fields_ = [("outN", C.c_void_p), ("outT", C.c_void_p), ("Rq", C.c_void_p), ("ms_n", C.c_double),
                ("ms_t", C.c_double)]


class EvalArgs(C.Structure):
    _fields_ = [("Y", C.c_void_p), ("y_dtype", C.c_int32), ("I", C.c_int64), ("J", C.c_int64),
                ("ldY", C.c_int64), ("y_storage", C.c_int32),
                ("Xu", C.c_void_p), ("Xv", C.c_void_p), ("Mx", C.c_int32), ("clamp8", C.c_int32),
                ("rowscale", C.c_void_p), ("colscale", C.c_void_p), ("beta_in", C.c_void_p),
                ("infer_beta", C.c_int32), ("device", C.c_int32)]


class EvalOut(C.Structure):
    _fields_ = [("alpha", C.c_void_p), ("beta", C.c_void_p), ("resid", C.c_void_p),
                ("sum_w", C.c_double), ("sum_ylogw", C.c_double), ("max_w", C.c_double),
                ("max_Y", C.c_double), ("LL", C.c_double)]


# every symbol include/scgbm.h declares
EXPORTS = [
    "scgbm_fit", "scgbm_fit_begin", "scgbm_fit_step", "scgbm_fit_profile", "scgbm_fit_set_profile", "scgbm_fit_end", "scgbm_get_se", "scgbm_project", "scgbm_sim_counts",
    "scgbm_sim_data", "scgbm_cci_perturb",
    "scgbm_dev_alloc", "scgbm_dev_free", "scgbm_memcpy_d2h", "scgbm_memcpy_h2d",
    "scgbm_host_alloc_pinned", "scgbm_host_free_pinned", "scgbm_flush_l2",
    "scgbm_comm_unique_id", "scgbm_comm_init_rank", "scgbm_comm_destroy", "scgbm_comm_allreduce_f64",
    "scgbm_last_error", "scgbm_device_count", "scgbm_version",
    "scgbm_dbg_tsvd", "scgbm_dbg_eval", "scgbm_dbg_eigh", "scgbm_dbg_prod", "scgbm_dbg_membench",
]

_lib = None


def load():
    """dlopen libscgbm.so.  Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScgbmError(2, f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()); "
                            "scgbm_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.scgbm_last_error.restype = C.c_char_p
    for name in EXPORTS:
        getattr(lib, name)   # AttributeError if the .so is stale
    lib.scgbm_fit.argtypes = [C.POINTER(FitArgs), C.POINTER(FitOut)]
    lib.scgbm_fit_begin.argtypes = [C.POINTER(FitArgs), C.POINTER(C.c_void_p)]
    lib.scgbm_fit_step.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
    lib.scgbm_fit_profile.argtypes = [C.c_void_p, C.POINTER(Profile)]
    lib.scgbm_fit_set_profile.argtypes = [C.c_void_p, C.c_int32]
    lib.scgbm_fit_end.argtypes = [C.c_void_p, C.POINTER(FitOut)]
    lib.scgbm_get_se.argtypes = [C.POINTER(SeArgs), C.POINTER(SeOut)]
    lib.scgbm_project.argtypes = [C.POINTER(ProjArgs), C.POINTER(ProjOut)]
    lib.scgbm_sim_counts.argtypes = [C.POINTER(SimArgs), C.c_void_p, C.POINTER(C.c_int64)]
    lib.scgbm_sim_data.argtypes = [C.POINTER(SimDataArgs), C.POINTER(SimDataOut)]
    lib.scgbm_cci_perturb.argtypes = [C.POINTER(CciArgs), C.c_void_p]
    lib.scgbm_dev_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_int]
    lib.scgbm_dev_free.argtypes = [C.c_void_p, C.c_int]
    lib.scgbm_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    lib.scgbm_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    lib.scgbm_host_alloc_pinned.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
    lib.scgbm_host_free_pinned.argtypes = [C.c_void_p]
    lib.scgbm_flush_l2.argtypes = [C.c_int]
    lib.scgbm_comm_unique_id.argtypes = [C.c_void_p]
    lib.scgbm_comm_init_rank.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.scgbm_comm_destroy.argtypes = [C.c_void_p]
    lib.scgbm_comm_allreduce_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    lib.scgbm_dbg_tsvd.argtypes = [C.POINTER(TsvdArgs), C.POINTER(TsvdOut)]
    lib.scgbm_dbg_prod.argtypes = [C.POINTER(ProdArgs), C.POINTER(ProdOut)]
    lib.scgbm_dbg_eval.argtypes = [C.POINTER(EvalArgs), C.POINTER(EvalOut)]
    lib.scgbm_dbg_eigh.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.scgbm_dbg_membench.argtypes = [C.c_int32, C.c_void_p]
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().scgbm_last_error()
        raise ScgbmError(status, (msg or b"unknown error").decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array (must stay alive while the call runs)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
