"""Diagnostic 2: where does the alpha error at full gene height come from -- eta precision (22-bit split) or the SVD?
alpha of iteration 18 belongs to the X produced by step 17 (kept factors of a max_iter = 17 run)."""
import ast
import os
import sys

import numpy as np
from scipy.special import logsumexp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import sim_counts_exact_rows, principal_angle  # noqa: E402
from scgbm_b200 import gbm_sc  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "tall_20000x2000_m20.npz"))
Y = sim_counts_exact_rows(**ast.literal_eval(str(G["gen"])))
rs = Y.sum(1).astype(np.float64)
n = int(G["max_iter"])
for tol in (3e-5, 1e-5, 1e-6):
    o17 = gbm_sc(Y, 20, max_iter=n - 1, return_W=False, diagnostics=True, keep_factors=True, svd_tol=tol)
    out = gbm_sc(Y, 20, max_iter=n, return_W=False, diagnostics=True, svd_tol=tol)
    X = o17["_Xu"] @ o17["_Xv"].T
    a_chk = np.log(rs) - logsumexp(o17["beta"][None, :] + X, axis=1)
    a_chk -= a_chk.mean()
    ea = np.abs(out["alpha"][:, 0] - G["alpha"][:, 0])
    ep = np.abs(out["alpha"][:, 0] - a_chk)
    es = np.abs(a_chk - G["alpha"][:, 0])
    mx = np.abs(X).max(1)
    print(f"svd_tol {tol:g}: passes {out['_prof']['svd_passes']} exact rows {out['_exact_rows']}  angle U {principal_angle(G['U'], out['U']):.2e} "
          f"V {principal_angle(G['V'], out['V']):.2e}\n  alpha vs oracle: max {ea.max():.2e}, #>1e-5: {(ea > 1e-5).sum()};  "
          f"eta-precision part: max {ep.max():.2e}, #>1e-5: {(ep > 1e-5).sum()};  SVD part (fp64 alpha of the device's X vs oracle): "
          f"max {es.max():.2e}, #>1e-5 {(es > 1e-5).sum()}\n  worst genes (SVD part): "
          f"{[(int(i), f'{es[i]:.1e}', f'max|x| {mx[i]:.0f}', int(rs[i])) for i in np.argsort(-es)[:6]]}", flush=True)
