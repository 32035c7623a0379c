"""Parity numbers of one synthetic case against the oracle (prints, no asserts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scgbm_b200 import gbm_sc
from oracle import scgbm_oracle as O
from tests.util import principal_angle, first_knife_edge

def case(I, J, d, M, kw, gen, **extra):
    Y = O.sim_data(I, J, d, seed=I + J, **gen)["Y"]
    ref = O.gbm_sc(Y, M, **kw)
    out = gbm_sc(Y, M, diagnostics=True, **kw, **extra)
    n = min(len(ref["_LL"]), len(out["_LL"]))
    rel = np.abs(ref["_LL"][:n] - out["_LL"][:n]) / np.abs(ref["_LL"][:n])
    print(f"{I}x{J} M={M} {kw} {extra}: iters {ref['_n_iter']}/{out['_n_iter']} knife {first_knife_edge(ref)} mask_eq "
          f"{np.array_equal(ref['_accepted'][:n], out['_accepted'][:n])} LLrel {rel.max():.2e} "
          f"alpha {np.abs(ref['alpha'] - out['alpha']).max():.2e} beta {np.abs(ref['beta'] - out['beta']).max():.2e} "
          f"angU {principal_angle(ref['U'], out['U']):.2e} angV {principal_angle(ref['V'], out['V']):.2e} "
          f"D {np.abs(ref['D'] - out['D']).max() / ref['D'][0]:.2e}", flush=True)

if __name__ == "__main__":
    for tol in (0.0, 1e-6):
        case(600, 1500, 10, 10, dict(), dict(nb_theta=1.0), svd_tol=tol)
        case(512, 640, 5, 20, dict(min_iter=10, max_iter=40), dict(nb_theta=2.0), svd_tol=tol)
        case(300, 400, 5, 5, dict(), dict(nb_theta=None), svd_tol=tol)
