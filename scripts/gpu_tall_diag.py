"""Diagnostic: the full-gene-height fixture (tests/golden/tall_20000x2000_m20.npz) against the CUDA path,
error by error, with and without the product rider."""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import sim_counts_exact_rows, principal_angle  # noqa: E402
from scgbm_b200 import gbm_sc  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "tall_20000x2000_m20.npz"))
Y = sim_counts_exact_rows(**ast.literal_eval(str(G["gen"])))
for rider in ("0", "force"):
    os.environ["SCGBM_FUSED_PROD"] = rider
    for extra in ({}, dict(svd_tol=1e-6)):
        out = gbm_sc(Y, 20, max_iter=int(G["max_iter"]), return_W=False, diagnostics=True, **extra)
        n = min(len(out["_LL"]), len(G["LL"]))
        rel = np.abs(out["_LL"][:n] - G["LL"][:n]) / np.abs(G["LL"][:n])
        ea = np.abs(out["alpha"] - G["alpha"])
        print(f"rider={rider} {extra}: LL rel max {rel.max():.2e} per-iter {np.array2string(rel, precision=1)}\n"
              f"  accepted equal {np.array_equal(out['_accepted'][:n], G['accepted'][:n])}  alpha err max {ea.max():.2e} "
              f"(median {np.median(ea):.2e}, argmax {ea.argmax()}, alpha there {G['alpha'].ravel()[ea.argmax()]:.3f}, "
              f"rowsum {Y[ea.argmax() % Y.shape[0]].sum()})  beta err {np.abs(out['beta'] - G['beta']).max():.2e}\n"
              f"  angle U {principal_angle(G['U'], out['U']):.2e} V {principal_angle(G['V'], out['V']):.2e} "
              f"D rel {np.abs(out['D'] / G['D'] - 1).max():.2e}  passes {out['_prof']['svd_passes']} "
              f"underflow reruns {out['_underflow_reruns']} unconverged {out['_svd_unconverged']}", flush=True)
        big = np.argsort(-ea.ravel())[:5]
        print("  worst genes:", [(int(i), f"{ea.ravel()[i]:.1e}", int(Y[i].sum())) for i in big])
