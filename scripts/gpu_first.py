"""First on-GPU check: eigh, tsvd, eval and the full fit against the oracle."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
from scgbm_b200 import _lib as L, gbm_sc
from oracle.r_rng import readme_demo_Y
from oracle import scgbm_oracle as O

lib = L.load()
print("devices", lib.scgbm_device_count())
rng = np.random.default_rng(0)
# eigh
for n in (5, 32, 44, 97, 128, 200):
    A = rng.standard_normal((n, n)); A = A @ A.T
    A0 = A.copy(); Af = np.asfortranarray(A); w = np.zeros(n)
    t = time.time(); L.check(lib.scgbm_dbg_eigh(L.ptr(Af), L.ptr(w), n, -1)); dt = time.time() - t
    wr = np.linalg.eigvalsh(A0)[::-1]
    res = np.abs(A0 @ Af - Af * w).max() / np.abs(wr).max()
    print(f"eigh n={n} val err {np.abs(w - wr).max() / wr.max():.2e} resid {res:.2e} orth {np.abs(Af.T @ Af - np.eye(n)).max():.2e} t={dt*1e3:.1f}ms")

# tsvd on a dense matrix
I, J, M = 700, 1900, 10
G = rng.standard_normal((I, 12)) @ np.diag(np.linspace(60, 20, 12)) @ rng.standard_normal((J, 12)).T / 30 + rng.standard_normal((I, J))
R32 = np.asfortranarray(G.astype(np.float32))
a = L.TsvdArgs(); a.I = I; a.J = J; a.R = L.ptr(R32); a.c = 1.0; a.nterms = 0; a.M = M; a.device = -1; a.svd_tol = 1e-6
U = np.zeros((I, M), order="F"); D = np.zeros(M); V = np.zeros((J, M), order="F")
o = L.TsvdOut(); o.U = L.ptr(U); o.D = L.ptr(D); o.V = L.ptr(V)
t = time.time(); L.check(lib.scgbm_dbg_tsvd(C.byref(a), C.byref(o))); dt = time.time() - t
ue, de, vte = np.linalg.svd(R32.astype(np.float64), full_matrices=False)
ang = np.arccos(np.clip(np.linalg.svd(ue[:, :M].T @ U, compute_uv=False).min(), -1, 1))
print(f"tsvd passes={o.passes} est={o.est:.2e} sv relerr {np.abs(D - de[:M]).max() / de[0]:.2e} angle {ang:.2e} t={dt:.3f}s")

# full fit README demo
Y = readme_demo_Y()
ref = O.gbm_sc(Y, 10)
t = time.time(); out = gbm_sc(Y, 10, diagnostics=True, profile=True); dt = time.time() - t
n = min(len(ref["_LL"]), len(out["_LL"]))
print("n_iter", ref["_n_iter"], out["_n_iter"], "mask equal", np.array_equal(ref["_accepted"], out["_accepted"]))
print("max rel LL diff", np.max(np.abs(ref["_LL"][:n] - out["_LL"][:n]) / np.abs(ref["_LL"][:n])))
print("alpha diff", np.abs(ref["alpha"] - out["alpha"]).max(), "beta diff", np.abs(ref["beta"] - out["beta"]).max())
for nm in ("U", "V"):
    s = np.linalg.svd(ref[nm].T @ out[nm], compute_uv=False)
    print(nm, "angle", np.arccos(np.clip(s.min(), -1, 1)))
print("D rel", np.abs(ref["D"] - out["D"]).max() / ref["D"][0], "W rel", np.abs(ref["W"] - out["W"]).max() / ref["W"].max())
print("fit time", dt, out["_prof"])
