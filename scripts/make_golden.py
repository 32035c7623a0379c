"""Generates the committed oracle vectors under tests/golden/ from the float64 oracle
(oracle/scgbm_oracle.py, itself pinned against the reference's README printout).  The CUDA path is
compared with these vectors in tests/test_gpu_golden.py / test_gpu_fit.py::test_golden_fixture with
no oracle fit at run time.  Every fixture is knife-edge free: if the oracle's run has an accept/revert
decision with a margin below 1e-7 (tests/util.py), the fixture is the run with max.iter shortened to
just before it, so alpha, beta, U, V, D and scores are always comparable.

    python scripts/make_golden.py [small] [c2] [tall]

  small  sim_nb_257x1031.npz      simData-recipe NB, M = 8, infer.beta = TRUE, + get.se  (Y stored)
  c2     c2_nb_2000x10000_m10.npz BASELINE configs[1]: 2k x 10k NB (theta = 1), M = 10, fit + get.se
  tall   tall_20000x2000_m20.npz  full gene height: 20 000 genes x 2 000 cells, ~88 % zeros, M = 20
The two large ones store only the fit; Y is regenerated from the seed (tests/util.py
sim_counts_exact_rows) and verified by its digest.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scgbm_oracle as O  # noqa: E402
from tests.util import first_knife_edge, sim_counts_exact_rows, digest  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def knife_free_fit(Y, M, **kw):
    ref = O.gbm_sc(Y, M, **kw)
    k = first_knife_edge(ref)
    if k < len(ref["_LL"]):
        kw = dict(kw, max_iter=k)
        ref = O.gbm_sc(Y, M, **kw)
        assert first_knife_edge(ref) == len(ref["_LL"])
    return ref, kw


def fit_fields(ref):
    return dict(LL=ref["_LL"], accepted=ref["_accepted"], obj=ref["obj"], n_iter=ref["_n_iter"],
                hit_max_iter=ref["_hit_max_iter"], alpha=ref["alpha"], beta=ref["beta"], U=ref["U"], V=ref["V"],
                D=ref["D"], scores=ref["scores"], alpha_xmax=ref["_alpha_xmax"])


def small():
    Y = O.sim_data(257, 1031, 8, seed=257 + 1031, nb_theta=1.0)["Y"]
    ref, kw = knife_free_fit(Y, 8, infer_beta=True)
    ref["W"] = ref["W"].astype(np.float32).astype(np.float64)     # stored as float32: get.se is evaluated on exactly this W
    ref = O.get_se_fast(ref)
    np.savez_compressed(os.path.join(GOLD, "sim_nb_257x1031.npz"), Y=Y.astype(np.int32), M=8, infer_beta=True,
                        max_iter=kw.get("max_iter", 100), W=ref["W"].astype(np.float32),
                        se_scores=ref["se_scores"], se_loadings=ref["se_loadings"], **fit_fields(ref))
    print("wrote sim_nb_257x1031.npz", Y.shape, ref["_n_iter"], "iterations, max_iter", kw.get("max_iter", 100))


def c2():
    gen = dict(I=2000, J=10000, d=10, seed=1, nb_theta=1.0)
    Y = sim_counts_exact_rows(**gen)
    t0 = time.time()
    ref, kw = knife_free_fit(Y, 10)
    ref = O.get_se_fast(ref)
    np.savez_compressed(os.path.join(GOLD, "c2_nb_2000x10000_m10.npz"), M=10, max_iter=kw.get("max_iter", 100),
                        y_digest=digest(Y), gen=np.array(repr(gen)), oracle_seconds=time.time() - t0,
                        se_scores=ref["se_scores"], se_loadings=ref["se_loadings"], **fit_fields(ref))
    print("wrote c2_nb_2000x10000_m10.npz", Y.shape, ref["_n_iter"], "iterations, max_iter", kw.get("max_iter", 100),
          f"{time.time() - t0:.0f} s")


def tall():
    gen = dict(I=20000, J=2000, d=20, seed=3, alpha_mean=-3.0, beta_mean=0.0)
    Y = sim_counts_exact_rows(**gen)
    t0 = time.time()
    ref, kw = knife_free_fit(Y, 20, max_iter=18)
    np.savez_compressed(os.path.join(GOLD, "tall_20000x2000_m20.npz"), M=20, max_iter=kw.get("max_iter", 18),
                        y_digest=digest(Y), gen=np.array(repr(gen)), zero_frac=float((Y == 0).mean()),
                        oracle_seconds=time.time() - t0, **fit_fields(ref))
    print("wrote tall_20000x2000_m20.npz", Y.shape, "zeros", (Y == 0).mean(), ref["_n_iter"], "iterations, max_iter",
          kw.get("max_iter", 18), f"{time.time() - t0:.0f} s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "c2", "tall"]
    for w in which:
        {"small": small, "c2": c2, "tall": tall}[w]()
