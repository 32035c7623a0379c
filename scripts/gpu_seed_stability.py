"""Does the END of a converging fit depend on rounding?  The 700 x 1800 multi-GPU test case on ONE GPU with different
seeds of the initial SVD's start block (the trajectory then differs by rounding only): objective and final factors
against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.scgbm_oracle import sim_data, gbm_sc as o_gbm_sc
from scgbm_b200 import gbm_sc
from tests.util import principal_angle
Y = sim_data(700, 1800, 8, seed=5, nb_theta=1.0)["Y"]
ref = o_gbm_sc(Y, 8, max_iter=40)
acc = ref["_accepted"].astype(bool)
print("oracle: n", len(ref["_LL"]), "rejected at", (np.nonzero(~acc)[0] + 1).tolist(), flush=True)
for seed in range(int(os.environ.get("NSEED", 8))):
    out = gbm_sc(Y, 8, diagnostics=True, max_iter=40, seed=seed)
    n = min(len(out["_LL"]), len(ref["_LL"]))
    rel = np.abs(out["_LL"][:n] - ref["_LL"][:n]) / np.abs(ref["_LL"][:n])
    same = np.array_equal(out["_accepted"][:n], ref["_accepted"][:n])
    print(f"seed {seed}: n_iter {out['_n_iter']} mask equal {same} | LL dev accepted {rel[acc[:n]].max():.2e} rejected {rel[~acc[:n]].max():.2e} (iter {int(np.argmax(rel)) + 1}) | "
          f"angle U {principal_angle(out['U'], ref['U']):.2e} V {principal_angle(out['V'], ref['V']):.2e} | passes {out['_prof']['svd_passes']} unconverged {out['_svd_unconverged']}", flush=True)
