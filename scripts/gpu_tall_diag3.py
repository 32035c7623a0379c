"""Diagnostic 3: pure eta-precision error of the row-sum pass (alpha) at full gene height: the device's alpha for a
given X (dbg_eval) against float64 on the same X."""
import ast, os, sys
import ctypes as C
import numpy as np
from scipy.special import logsumexp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import sim_counts_exact_rows  # noqa: E402
from scgbm_b200 import gbm_sc, _lib as L  # noqa: E402
G = np.load(os.path.join(ROOT, "tests", "golden", "tall_20000x2000_m20.npz"))
Y = sim_counts_exact_rows(**ast.literal_eval(str(G["gen"])))
out = gbm_sc(Y, 20, max_iter=17, return_W=False, diagnostics=True, keep_factors=True, svd_tol=1e-6)
Xu, Xv = np.asfortranarray(out["_Xu"]), np.asfortranarray(out["_Xv"])
X = Xu @ Xv.T
beta = np.ascontiguousarray(out["beta"])
lib = L.load()
I, J = Y.shape
a = L.EvalArgs(); Yh = np.asfortranarray(Y.astype(np.int32))
a.Y = L.ptr(Yh); a.y_dtype = L.I32; a.I = I; a.J = J; a.ldY = I; a.y_storage = 0
a.Xu = L.ptr(Xu); a.Xv = L.ptr(Xv); a.Mx = 20; a.clamp8 = 0; a.beta_in = L.ptr(beta); a.infer_beta = 0; a.device = -1
al = np.zeros(I); be = np.zeros(J)
o = L.EvalOut(); o.alpha = L.ptr(al); o.beta = L.ptr(be)
L.check(lib.scgbm_dbg_eval(C.byref(a), C.byref(o)))
rs = Y.sum(1).astype(np.float64)
a_ref = np.log(rs) - logsumexp(beta[None, :] + X, axis=1)
a_ref -= a_ref.mean()
err = np.abs(al - a_ref)
mx = np.abs(X).max(1)
print(f"eta-precision alpha error: max {err.max():.2e}, #>1e-5 {(err > 1e-5).sum()}, #>3e-6 {(err > 3e-6).sum()}")
for lo, hi in ((0, 4), (4, 8), (8, 16), (16, 32), (32, 100)):
    m = (mx >= lo) & (mx < hi)
    if m.any():
        print(f"  rows with max|X| in [{lo},{hi}): {m.sum():6d}  alpha err max {err[m].max():.2e} median {np.median(err[m]):.2e}")
rn = np.linalg.norm(Xu, axis=1); bn = np.linalg.norm(Xv, axis=1).max()
print(f"row-norm bound |a_i| max|b_j|: rows > 8: {(rn * bn > 8).sum()}, > 16: {(rn * bn > 16).sum()}, > 32: {(rn * bn > 32).sum()}; max|b| {bn:.3f}")
