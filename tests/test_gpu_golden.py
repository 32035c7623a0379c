"""CUDA path against the committed oracle fixtures of the two named shapes the oracle cannot be run on
inside a GPU test (scripts/make_golden.py; the oracle needs minutes there):

  * BASELINE configs[1] ("C2"): simData-style 2k genes x 10k cells NB counts, M = 10, full fit + get.se;
  * full gene height: 20 000 genes x 2 000 cells (~88 % zeros), M = 20, 18 iterations against the
    exact-SVD oracle -- 313 gene tiles per cell tile, the 157-drain cadence of the product rider and the
    fp32 chunk sums at I = 20 000 are exactly what the small parity cases cannot reach.

Y is regenerated from the seed (tests/util.py) and verified by its digest; tolerances are north_star's."""
import ast
import os

import numpy as np
import pytest

from tests.util import assert_fit_parity, digest, sim_counts_exact_rows

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _fixture(name):
    G = np.load(os.path.join(GOLD, name))
    ref = {k: G[k] for k in ("alpha", "beta", "U", "V", "D", "scores", "obj")}
    ref["_LL"] = G["LL"]; ref["_accepted"] = G["accepted"]; ref["_n_iter"] = int(G["n_iter"])
    ref["_alpha_xmax"] = G["alpha_xmax"]
    Y = sim_counts_exact_rows(**ast.literal_eval(str(G["gen"])))
    assert digest(Y) == str(G["y_digest"]), "the generator no longer reproduces the fixture's count matrix"
    return G, ref, Y


def test_c2_fit_and_get_se_match_oracle_fixture():
    from scgbm_b200 import gbm_sc, get_se
    from oracle.scgbm_oracle import get_se_fast
    G, ref, Y = _fixture("c2_nb_2000x10000_m10.npz")
    assert Y.shape == (2000, 10000)
    out = gbm_sc(Y, 10, max_iter=int(G["max_iter"]), diagnostics=True)
    assert assert_fit_parity(ref, out, check_W=False) == "full"
    assert bool(out["_hit_max_iter"]) == bool(G["hit_max_iter"])
    se = get_se(out)
    # (1) the get.se kernels on the CUDA fit's own (U, scores, W): oracle arithmetic on identical inputs
    mine = get_se_fast(out)
    np.testing.assert_allclose(se["se_scores"], mine["se_scores"], rtol=1e-9)
    np.testing.assert_allclose(se["se_loadings"], mine["se_loadings"], rtol=1e-9)
    # (2) end to end against the ORACLE fit's standard errors: the fits agree to the parity tolerances, the
    # simData spectrum is separated (10 % gaps), so the per-factor SEs agree to ~1e-4
    np.testing.assert_allclose(se["se_scores"], G["se_scores"], rtol=2e-3)
    np.testing.assert_allclose(se["se_loadings"], G["se_loadings"], rtol=2e-3)
    assert np.median(np.abs(se["se_scores"] / G["se_scores"] - 1)) < 1e-4


@pytest.mark.parametrize("rider", ["0", "force"])
def test_full_gene_height_matches_oracle_fixture(rider, monkeypatch):
    from scgbm_b200 import gbm_sc
    G, ref, Y = _fixture("tall_20000x2000_m20.npz")
    assert Y.shape == (20000, 2000) and int(G["n_iter"]) >= 15
    monkeypatch.setenv("SCGBM_FUSED_PROD", rider)
    out = gbm_sc(Y, 20, max_iter=int(G["max_iter"]), return_W=False, diagnostics=True)
    assert assert_fit_parity(ref, out, check_W=False) == "full"
    if rider == "force":
        monkeypatch.setenv("SCGBM_FUSED_PROD", "0")
        plain = gbm_sc(Y, 20, max_iter=int(G["max_iter"]), return_W=False, diagnostics=True)
        assert out["_prof"]["svd_passes"] < plain["_prof"]["svd_passes"]      # the rider really ran at full height
