"""Generates tests/golden/sim_nb_257x1031.npz from the float64 oracle (oracle/scgbm_oracle.py,
itself pinned against the reference's README printout): a seeded simData-recipe NB matrix, the
oracle's gbm.sc(Y, M=8, infer.beta=TRUE) fit and get.se.  The CUDA path is compared with these
committed vectors in tests/test_gpu_fit.py::test_golden_fixture (no oracle needed at run time).

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scgbm_oracle as O  # noqa: E402

Y = O.sim_data(257, 1031, 8, seed=257 + 1031, nb_theta=1.0)["Y"]
ref = O.gbm_sc(Y, 8, infer_beta=True)
ref["W"] = ref["W"].astype(np.float32).astype(np.float64)     # stored as float32: get.se is evaluated on exactly this W
ref = O.get_se_fast(ref)
np.savez_compressed(
    os.path.join(ROOT, "tests", "golden", "sim_nb_257x1031.npz"),
    Y=Y.astype(np.int32), M=8, infer_beta=True,
    LL=ref["_LL"], accepted=ref["_accepted"], obj=ref["obj"], n_iter=ref["_n_iter"],
    alpha=ref["alpha"], beta=ref["beta"], U=ref["U"], V=ref["V"], D=ref["D"], scores=ref["scores"],
    W=ref["W"].astype(np.float32), se_scores=ref["se_scores"], se_loadings=ref["se_loadings"])
print("wrote tests/golden/sim_nb_257x1031.npz", Y.shape, ref["_n_iter"], "iterations")
