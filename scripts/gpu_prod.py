"""On-GPU check of the K5 residual products (FFMA vs tcgen05) against float64 numpy."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
from scgbm_b200 import _lib as L

lib = L.load()


def run(I, J, b, variant, reps=1, seed=0, check=True):
    rng = np.random.default_rng(seed)
    R = np.asfortranarray(rng.standard_normal((I, J)).astype(np.float32) * (1 + 3 * rng.random((I, 1)).astype(np.float32)))
    P = np.asfortranarray(rng.standard_normal((J, b)) / np.sqrt(J))
    Q = np.asfortranarray(rng.standard_normal((I, b)) / np.sqrt(I))
    a = L.ProdArgs(); a.I = I; a.J = J; a.R = L.ptr(R); a.P = L.ptr(P); a.Q = L.ptr(Q); a.b = b; a.c = 0.37
    a.variant = variant; a.reps = reps; a.device = -1
    oN = np.zeros((I, b), order="F"); oT = np.zeros((J, b), order="F")
    Rq = np.zeros((I, J), order="F") if check else None
    o = L.ProdOut(); o.outN = L.ptr(oN); o.outT = L.ptr(oT); o.Rq = L.ptr(Rq)
    L.check(lib.scgbm_dbg_prod(C.byref(a), C.byref(o)))
    res = {"ms_n": o.ms_n, "ms_t": o.ms_t}
    if check:
        refN = 0.37 * (Rq @ P); refT = 0.37 * (Rq.T @ Q)
        res["errN"] = float(np.abs(oN - refN).max() / np.abs(refN).max())
        res["errT"] = float(np.abs(oT - refT).max() / np.abs(refT).max())
        res["quant"] = float(np.abs(Rq - R).max() / np.abs(R).max())
    return res


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "small"):
        for (I, J, b) in [(128, 64, 32), (128, 512, 32), (256, 1024, 32), (500, 500, 22), (300, 421, 17), (2000, 10000, 22), (640, 3000, 62)]:
            for v in (1, 2):
                print(I, J, b, "variant", v, run(I, J, b, v), flush=True)
    if which in ("all", "big"):
        I, J, b = 20000, 100000, 32
        for v in (1, 2):
            r = run(I, J, b, v, reps=5, check=False)
            gb = I * J * 4 / 1e9
            print("big", I, J, b, "variant", v, r, f"GB/s n={gb / r['ms_n'] * 1e3:.0f} t={gb / r['ms_t'] * 1e3:.0f}", flush=True)
