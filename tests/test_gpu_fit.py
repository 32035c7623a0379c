"""Parity of the CUDA gbm.sc / get.se / projection path (through the C ABI) with the
float64 oracle, on the reference's own demo and on seeded synthetic inputs."""
import json
import os
import warnings

import numpy as np
import pytest

from tests.util import assert_fit_parity, checked_fit, principal_angle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sim(I, J, d, seed, nb=None, amean=0.0, bmean=0.0):
    from oracle.scgbm_oracle import sim_data
    return sim_data(I, J, d, alpha_mean=amean, beta_mean=bmean, seed=seed, nb_theta=nb)["Y"]


def test_readme_demo_matches_reference_printout(readme_Y):
    """The reference's only golden vector (README.md:36-62): the CUDA path must print the
    same 27 objectives, skip iterations 12 and 19, and stop after 29."""
    from scgbm_b200 import gbm_sc
    G = json.load(open(os.path.join(GOLD, "readme_demo.json")))
    out = gbm_sc(readme_Y, 10, diagnostics=True)
    assert (np.nonzero(out["_accepted"])[0] + 1).tolist() == G["printed_iterations"]
    assert ["%.7g" % v for v in out["obj"]] == G["printed_objective"]
    assert out["_n_iter"] == 31


def test_readme_demo_full_parity(readme_Y, readme_oracle_fit):
    from scgbm_b200 import gbm_sc
    out = gbm_sc(readme_Y, 10, diagnostics=True)
    assert assert_fit_parity(readme_oracle_fit, out) == "full"
    # list layout (quirk Q13) and Q7: loadings keep the pre-flip signs
    assert list(k for k in out if not k.startswith("_")) == \
        ["W", "V", "D", "U", "loadings", "alpha", "beta", "M", "I", "J", "obj", "scores"]
    flip = np.sign(out["loadings"][0, :])
    flip[flip == 0] = 1
    np.testing.assert_allclose(out["loadings"] * flip[None, :], out["U"], rtol=0, atol=0)


def test_readme_get_se(readme_Y, readme_oracle_fit):
    from scgbm_b200 import gbm_sc, get_se
    from oracle.scgbm_oracle import get_se as o_get_se
    G = json.load(open(os.path.join(GOLD, "readme_demo.json")))
    out = get_se(gbm_sc(readme_Y, 10))
    np.testing.assert_allclose(out["se_scores"][:6], np.array(G["se_scores_head"]), rtol=0, atol=5e-5)
    # same inputs -> same standard errors as the oracle to fp64 accuracy
    ref = o_get_se(readme_oracle_fit)
    mine = get_se(readme_oracle_fit)
    np.testing.assert_allclose(mine["se_scores"], ref["se_scores"], rtol=1e-10)
    np.testing.assert_allclose(mine["se_loadings"], ref["se_loadings"], rtol=1e-10)
    eps = get_se(readme_oracle_fit, EPS=0.001)
    np.testing.assert_allclose(eps["se_scores"], o_get_se(readme_oracle_fit, EPS=0.001)["se_scores"], rtol=1e-10)


def test_golden_fixture():
    """CUDA path against committed oracle vectors (scripts/make_golden.py): no oracle at run time."""
    from scgbm_b200 import gbm_sc, get_se
    G = np.load(os.path.join(GOLD, "sim_nb_257x1031.npz"))
    out = gbm_sc(G["Y"], int(G["M"]), infer_beta=bool(G["infer_beta"]), max_iter=int(G["max_iter"]), diagnostics=True)
    ref = {k: G[k] for k in ("alpha", "beta", "U", "V", "D", "scores", "obj")}
    ref["_LL"] = G["LL"]; ref["_accepted"] = G["accepted"]; ref["_n_iter"] = int(G["n_iter"])
    ref["_alpha_xmax"] = G["alpha_xmax"]
    # the fixture is knife-edge free (scripts/make_golden.py): alpha, beta, U, V, D, scores are all compared
    assert assert_fit_parity(ref, out, check_W=False) == "full"
    assert bool(out["_hit_max_iter"]) == bool(G["hit_max_iter"])
    # get.se on the fixture's own fit (per-factor SEs are only comparable for identical U, scores, W)
    se = get_se({"U": G["U"], "scores": G["scores"], "D": G["D"], "W": G["W"].astype(np.float64)})
    np.testing.assert_allclose(se["se_scores"], G["se_scores"], rtol=1e-8)
    np.testing.assert_allclose(se["se_loadings"], G["se_loadings"], rtol=1e-8)


CASES = [
    # I, J, d(sim), M, kwargs, generator kwargs
    (300, 400, 5, 5, dict(), dict(nb=None)),
    (257, 1031, 8, 8, dict(infer_beta=True), dict(nb=1.0)),
    (600, 1500, 10, 10, dict(), dict(nb=1.0)),
    (150, 220, 3, 1, dict(), dict()),                                   # M = 1 (NEWS.md:33)
    (400, 300, 4, 6, dict(max_iter=12), dict()),                        # hits max.iter
    (333, 777, 6, 6, dict(tol=1e-10, infer_beta=True, max_iter=60), dict(amean=-1.0)),   # figures/scGBM_sim.R:85
    (512, 640, 5, 20, dict(min_iter=10, max_iter=40), dict(nb=2.0)),    # M = 20
    (400, 520, 6, 50, dict(max_iter=8), dict()),                        # M = 50 (config 5): 64-wide Krylov block, 3 K slabs
    (40, 33, 2, 2, dict(max_iter=15), dict()),                          # smaller than one tile in both directions
]


@pytest.mark.parametrize("I,J,d,M,kw,gen", CASES)
def test_fit_matches_oracle(I, J, d, M, kw, gen):
    Y = _sim(I, J, d, seed=I + J, **gen)
    with warnings.catch_warnings(record=True) as wl:
        warnings.simplefilter("always")
        ref, out, info = checked_fit(Y, M, **kw)      # full parity, also when the requested run has a knife edge
    assert any("Maximum number of iterations" in str(w.message) for w in wl) == \
        (bool(out["_hit_max_iter"]) or bool(info["requested"][1]["_hit_max_iter"]))
    # invariants
    assert abs(out["alpha"].mean()) < 1e-10
    assert np.all(np.diff(out["D"]) <= 1e-9 * out["D"][0])
    assert np.abs(out["U"].T @ out["U"] - np.eye(M)).max() < 1e-8
    assert np.abs(out["V"].T @ out["V"] - np.eye(M)).max() < 1e-6


@pytest.mark.parametrize("storage,dtype", [(1, np.int32), (2, np.int32), (3, np.int32), (0, np.float64), (0, np.uint8), (0, np.float32)])
def test_storage_formats_agree(storage, dtype):
    """u8 / u16 / f32 device formats and every host dtype give the same fit."""
    from scgbm_b200 import gbm_sc
    Y = np.minimum(_sim(200, 300, 4, seed=11), 255)
    base = gbm_sc(Y.astype(np.int32), 4, diagnostics=True, max_iter=15, y_storage=2)
    out = gbm_sc(Y.astype(dtype), 4, diagnostics=True, max_iter=15, y_storage=storage)
    # u8 / u16 run the same tcgen05 kernels -> identical sums; f32 storage takes the CUDA-core kernels
    np.testing.assert_allclose(out["_LL"], base["_LL"], rtol=1e-12 if storage != 3 else 2e-7)
    assert out["_y_storage"] == (storage if storage else 1)      # AUTO: max(Y) <= 255 here -> u8 (s_Y = 1)
    if storage == 0 and dtype is np.float64:
        big = Y.astype(np.int32).copy(); big[0, 0] = 300              # one count above 255: AUTO stays at u16
        assert gbm_sc(big, 4, diagnostics=True, max_iter=3)["_y_storage"] == 2


def test_get_se_without_W():
    """get.se after return.W = FALSE, W re-evaluated on the device from the kept factors
    (Q6: the reverted X) -- same standard errors as from the returned W."""
    from scgbm_b200 import gbm_sc, get_se
    rng = np.random.default_rng(5)
    Y = _sim(240, 610, 5, seed=31, nb=1.0)
    full = gbm_sc(Y, 5, keep_factors=True)
    assert full["_x_rank"] == 5
    se_w = get_se(full)
    lean = dict(full); lean["W"] = None
    se_f = get_se(lean)
    np.testing.assert_allclose(se_f["se_scores"], se_w["se_scores"], rtol=1e-9)
    np.testing.assert_allclose(se_f["se_loadings"], se_w["se_loadings"], rtol=1e-9)
    # with batches
    labels = rng.integers(0, 2, size=Y.shape[1])
    Yb = Y + 1 * (rng.random(Y.shape) < 0.2)
    fb = gbm_sc(Yb, 4, batch=labels, max_iter=12, keep_factors=True)
    lean = dict(fb); lean["W"] = None
    np.testing.assert_allclose(get_se(lean)["se_scores"], get_se(fb)["se_scores"], rtol=1e-9)
    # no W and no factors: the reference's own failure mode (Q9)
    from scgbm_b200.api import RError
    with pytest.raises(RError, match="needs out"):
        get_se(gbm_sc(Y, 5, return_W=False, max_iter=5))


def test_sparse_input_equals_dense():
    """dgCMatrix-style column-compressed input gives bit-identical sums to the dense path."""
    import scipy.sparse as sp
    from scgbm_b200 import gbm_sc
    from scgbm_b200.api import RError
    Y = _sim(311, 777, 5, seed=21, amean=-1.0, bmean=-1.0)
    dense = gbm_sc(Y, 5, diagnostics=True, max_iter=12)
    for S in (sp.csc_matrix(Y.astype(np.float64)), sp.csr_matrix(Y.astype(np.int32))):
        out = gbm_sc(S, 5, diagnostics=True, max_iter=12)
        np.testing.assert_allclose(out["_LL"], dense["_LL"], rtol=1e-12)
        np.testing.assert_allclose(out["alpha"], dense["alpha"], atol=1e-12)
        assert out["_max_Y"] == dense["_max_Y"]
    bad = sp.csc_matrix(Y.astype(np.float64)); bad.data[3] = -1.0
    with pytest.raises(RError, match="negative"):
        gbm_sc(bad, 5)


def test_non_integer_counts_fall_back_to_f32():
    from scgbm_b200 import gbm_sc
    from oracle.scgbm_oracle import gbm_sc as o_gbm_sc
    Y = _sim(120, 160, 3, seed=3).astype(np.float64) * 0.5
    out = gbm_sc(Y, 3, diagnostics=True, max_iter=10)
    assert out["_y_storage"] == 3
    ref = o_gbm_sc(Y, 3, max_iter=10)
    np.testing.assert_allclose(out["_LL"], ref["_LL"], rtol=1e-6)


def test_batches_match_oracle():
    from scgbm_b200 import gbm_sc
    from oracle.scgbm_oracle import gbm_sc as o_gbm_sc
    rng = np.random.default_rng(4)
    Y = _sim(220, 500, 4, seed=8)
    J = Y.shape[1]
    labels = rng.choice(np.array(["b", "a", "c"]), size=J)
    Y = Y + 1 * (rng.random(Y.shape) < 0.2)       # every gene has counts in every batch
    codes = np.unique(labels, return_inverse=True)[1] + 1
    for kw in (dict(), dict(infer_beta=True)):
        ref, out, _ = checked_fit(Y, 4, cuda_kw=dict(batch=labels), oracle_kw=dict(batch=codes), max_iter=25, **kw)
        assert out["alpha"].shape == (Y.shape[0], 3)


def test_input_validation_mirrors_reference():
    """R/fit.R:287-299 and the Q11 precondition."""
    from scgbm_b200 import gbm_sc, RError
    Y = _sim(50, 60, 2, seed=1).astype(np.float64)
    bad = Y.copy(); bad[3, 4] = np.nan
    with pytest.raises(RError, match="NA values"):
        gbm_sc(bad, 2)
    bad = Y.copy(); bad[3, 4] = -1
    with pytest.raises(RError, match="negative"):
        gbm_sc(bad, 2)
    with pytest.raises(RError, match="M must be >= 1"):
        gbm_sc(Y, 0)
    bad = Y.copy(); bad[7, :] = 0
    with pytest.raises(RError, match="at least one count"):
        gbm_sc(bad, 2)
    with pytest.raises(RError, match="Batch"):
        gbm_sc(Y, 2, batch=np.zeros(3))


def test_return_W_false_and_time_by_iter():
    from scgbm_b200 import gbm_sc, get_se, RError
    Y = _sim(100, 140, 3, seed=2)
    out = gbm_sc(Y, 3, return_W=False, time_by_iter=True, max_iter=8)
    assert "W" not in out and "I" not in out and "J" not in out      # quirk Q9
    assert len(out["time"]) == 8 and np.all(np.diff(out["time"]) > 0)
    with pytest.raises(RError):
        get_se(out)


def test_get_se_matches_oracle_on_fit():
    from scgbm_b200 import gbm_sc, get_se
    from oracle.scgbm_oracle import get_se_fast
    Y = _sim(310, 450, 5, seed=21, nb=1.0)
    out = get_se(gbm_sc(Y, 5))
    ref = get_se_fast(out)
    np.testing.assert_allclose(out["se_scores"], ref["se_scores"], rtol=1e-9)
    np.testing.assert_allclose(out["se_loadings"], ref["se_loadings"], rtol=1e-9)
    assert out["se_scores"].shape == (Y.shape[1], 5) and out["se_loadings"].shape == (Y.shape[0], 5)


def test_projection_matches_oracle():
    """gbm.sc(subset = idx)  (R/fit.R:213-285) with an explicit index vector.  The
    subsample fit is deterministic, so the oracle's projection is run on the SAME
    subsample fit: every per-cell GLM, the re-centring and the re-expansion must agree."""
    from scgbm_b200 import gbm_sc
    from oracle.scgbm_oracle import gbm_proj_parallel as o_proj, gbm_sc as o_gbm_sc
    Y = _sim(260, 900, 4, seed=31, bmean=0.5)
    idx = np.arange(1, 301)                       # 1-based like R
    out = gbm_sc(Y, 4, subset=idx)
    Ysub = Y[:, idx - 1]
    ixs = np.nonzero(Ysub.sum(1) > 5)[0]
    sub = gbm_sc(Ysub[ixs], 4, diagnostics=True)
    ref = o_proj(Y.astype(np.float64), 4, idx, _subfit=sub)
    np.testing.assert_array_equal(ref["_ixs"], out["_ixs"])
    assert not out["_fail"].any() and not ref["_fail"].any()
    np.testing.assert_allclose(out["scores"], ref["scores"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(out["se_scores"], ref["se_scores"], rtol=1e-7)
    np.testing.assert_allclose(out["beta"], ref["beta"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out["alpha"], ref["alpha"], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(out["U"], ref["U"])
    assert out["V"] is None and out["scores"].shape == (Y.shape[1], 4)
    assert np.all(out["U"][np.setdiff1d(np.arange(Y.shape[0]), out["_ixs"])] == 0)
    # and the subsample fit itself is a parity-checked fit
    checked_fit(Ysub[ixs], 4)
    # sparse (dgCMatrix-style) input takes the same path after the on-device scatter
    import scipy.sparse as sp
    outs = gbm_sc(sp.csc_matrix(Y.astype(np.float64)), 4, subset=idx)
    np.testing.assert_allclose(outs["scores"], out["scores"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(outs["se_scores"], out["se_scores"], rtol=1e-12)
    np.testing.assert_allclose(outs["alpha"], out["alpha"], rtol=0, atol=1e-12)


def test_projection_glm_kernel_exact():
    """The per-cell IRLS kernel against the fastglm restatement on the SAME design."""
    import ctypes as C
    from scgbm_b200 import _lib as L
    from oracle.scgbm_oracle import poisson_glm_irls
    rng = np.random.default_rng(77)
    Ik, J, M = 180, 37, 6
    U, _ = np.linalg.qr(rng.standard_normal((Ik, M)))
    alpha = rng.normal(0, 0.7, Ik)
    sc = rng.normal(0, 3.0, (J, M)); b0 = rng.normal(0.2, 0.5, J)
    Y = rng.poisson(np.exp(alpha[:, None] + b0[None, :] + U @ sc.T)).astype(np.int32)
    Yf = np.asfortranarray(Y)
    ixs = np.arange(Ik, dtype=np.int64)
    beta = np.zeros(J); scores = np.zeros((J, M), order="F"); se = np.zeros((J, M), order="F")
    fail = np.zeros(J, dtype=np.int8); iters = np.zeros(J, dtype=np.int32)
    Uf = np.asfortranarray(U)
    a = L.ProjArgs()
    a.Y = L.ptr(Yf); a.y_dtype = L.I32; a.I_full = Ik; a.J = J; a.ldY = Ik; a.ixs = L.ptr(ixs); a.Ik = Ik; a.M = M
    a.U = L.ptr(Uf); a.alpha = L.ptr(alpha); a.glm_tol = 1e-8; a.glm_maxit = 100; a.device = -1
    o = L.ProjOut(); o.beta = L.ptr(beta); o.scores = L.ptr(scores); o.se_scores = L.ptr(se)
    o.fail = L.ptr(fail); o.iters = L.ptr(iters)
    L.check(L.load().scgbm_project(C.byref(a), C.byref(o)))
    Xd = np.hstack([np.ones((Ik, 1)), U])
    for j in range(J):
        coef, sej, conv, nit = poisson_glm_irls(Xd, Y[:, j].astype(float), alpha)
        assert iters[j] == nit
        np.testing.assert_allclose(beta[j], coef[0], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(scores[j], coef[1:], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(se[j], sej[1:], rtol=1e-9)
    assert not fail.any()


def test_stepwise_handle_equals_one_shot():
    """scgbm_fit_begin/_step/_end is the same fit as scgbm_fit."""
    import ctypes as C
    from scgbm_b200 import _lib as L, gbm_sc
    Y = np.asfortranarray(_sim(150, 200, 3, seed=6).astype(np.int32))
    I, J = Y.shape; M = 3
    ref = gbm_sc(Y, M, diagnostics=True)
    lib = L.load()
    a = L.FitArgs(); a.Y = L.ptr(Y); a.y_dtype = L.I32; a.I = I; a.J = J; a.ldY = I; a.M = M
    a.max_iter = 100; a.tol = 1e-4; a.min_iter = 30; a.nbatch = 1; a.device = -1; a.return_W = 0
    h = C.c_void_p()
    L.check(lib.scgbm_fit_begin(C.byref(a), C.byref(h)))
    done = C.c_int32(0); trips = 0
    while not done.value:
        L.check(lib.scgbm_fit_step(h, 1, C.byref(done))); trips += 1
    U = np.zeros((I, M), order="F"); V = np.zeros((J, M), order="F"); D = np.zeros(M)
    alpha = np.zeros(I); beta = np.zeros(J); obj = np.zeros(100)
    o = L.FitOut(); o.U = L.ptr(U); o.V = L.ptr(V); o.D = L.ptr(D); o.alpha = L.ptr(alpha); o.beta = L.ptr(beta); o.obj = L.ptr(obj)
    L.check(lib.scgbm_fit_end(h, C.byref(o)))
    assert o.n_iter == ref["_n_iter"] == trips
    np.testing.assert_array_equal(obj[: o.n_obj], ref["obj"])
    np.testing.assert_array_equal(U, ref["U"])


@pytest.mark.parametrize("I,J,d,M,kw,gen", [CASES[2], CASES[6], (257, 1031, 8, 8, dict(infer_beta=True), dict(nb=1.0))])
def test_fused_product_rider_matches_plain_pass(I, J, d, M, kw, gen, monkeypatch):
    """The fused pass's rider (R^T Q0 from the residual tiles in shared memory; eta_tc.cu PROD) replaces the
    warm SVD's first pass over the planes: same tiles, same operand planes, same drain cadence as prod_t."""
    from scgbm_b200 import gbm_sc
    from oracle.scgbm_oracle import gbm_sc as o_gbm_sc
    from tests.util import principal_angle
    Y = _sim(I, J, d, seed=I + J, **gen)
    monkeypatch.setenv("SCGBM_FUSED_PROD", "0")
    plain = gbm_sc(Y, M, diagnostics=True, **kw)
    monkeypatch.setenv("SCGBM_FUSED_PROD", "force")
    rider = gbm_sc(Y, M, diagnostics=True, **kw)
    assert rider["_prof"]["svd_passes"] < plain["_prof"]["svd_passes"]          # the rider really ran
    n = min(len(plain["_LL"]), len(rider["_LL"]))
    np.testing.assert_array_equal(plain["_accepted"][:n], rider["_accepted"][:n])
    np.testing.assert_array_equal(rider["_LL"][:n], plain["_LL"][:n])     # bit-identical: see eta_tc.cu PROD
    np.testing.assert_array_equal(rider["alpha"], plain["alpha"])
    assert principal_angle(rider["U"], plain["U"]) < 1e-12
    checked_fit(Y, M, **kw)          # still under SCGBM_FUSED_PROD=force


@pytest.mark.parametrize("variant", ["batches", "u8"])
def test_fused_product_rider_other_instantiations(variant, monkeypatch):
    """The rider's multi-batch and u8-count instantiations; the whole fit is bit-identical with and without it
    (the rider's R^T Q0 equals prod_t's bit for bit, and c * (2^k acc) rounds like (c 2^k) * acc)."""
    from scgbm_b200 import gbm_sc
    rng = np.random.default_rng(11)
    Y = _sim(300, 700, 5, seed=21)
    kw = dict(max_iter=20)
    if variant == "batches":
        Y = Y + 1 * (rng.random(Y.shape) < 0.2)
        kw["batch"] = rng.choice(np.array(["x", "y"]), size=Y.shape[1])
    else:
        Y = np.minimum(Y, 200)
        kw["y_storage"] = 1        # SCGBM_STORE_U8
    monkeypatch.setenv("SCGBM_FUSED_PROD", "0")
    plain = gbm_sc(Y, 5, diagnostics=True, **kw)
    monkeypatch.setenv("SCGBM_FUSED_PROD", "force")
    rider = gbm_sc(Y, 5, diagnostics=True, **kw)
    assert rider["_prof"]["svd_passes"] < plain["_prof"]["svd_passes"]
    np.testing.assert_array_equal(rider["_LL"], plain["_LL"])
    np.testing.assert_array_equal(rider["U"], plain["U"])
    np.testing.assert_array_equal(rider["V"], plain["V"])


@pytest.mark.parametrize("rider", ["0", "force"])
def test_packed_epilogue_agrees_with_the_plain_form(rider, monkeypatch):
    """SCGBM_FUSED_PACKED=1 selects the packed-fp32 epilogue of the fused pass (FADD2 / FMUL2 / FFMA2; eta_tc.cu PK).
    It rounds the residual differently in the last bit, nothing more: same accept / revert decisions, objective
    within 1e-8 relative, alpha within 1e-7 -- far inside the parity tolerances; the default is the plain form."""
    from scgbm_b200 import gbm_sc
    from tests.util import principal_angle
    Y = _sim(300, 700, 5, seed=21)
    monkeypatch.setenv("SCGBM_FUSED_PROD", rider)
    monkeypatch.setenv("SCGBM_FUSED_PACKED", "0")
    plain = gbm_sc(Y, 5, diagnostics=True, max_iter=20)
    monkeypatch.setenv("SCGBM_FUSED_PACKED", "1")
    packed = gbm_sc(Y, 5, diagnostics=True, max_iter=20)
    np.testing.assert_array_equal(packed["_accepted"], plain["_accepted"])
    assert not np.array_equal(packed["_LL"], plain["_LL"])       # the knob really switched the kernel
    np.testing.assert_allclose(packed["_LL"], plain["_LL"], rtol=1e-8)
    np.testing.assert_allclose(packed["alpha"], plain["alpha"], atol=1e-7)
    assert principal_angle(packed["U"], plain["U"]) < 1e-6


def _glm_problem(Ik, J, M, seed, scale=3.0):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((Ik, M)))
    alpha = rng.normal(0, 0.7, Ik)
    sc = rng.normal(0, scale, (J, M)); b0 = rng.normal(0.2, 0.5, J)
    Y = rng.poisson(np.exp(alpha[:, None] + b0[None, :] + U @ sc.T)).astype(np.int32)
    return U, alpha, Y


def _project(U, alpha, Y, method, tol=1e-8):
    import ctypes as C
    from scgbm_b200 import _lib as L
    Ik, J = Y.shape; M = U.shape[1]
    Yf = np.asfortranarray(Y); Uf = np.asfortranarray(U); ixs = np.arange(Ik, dtype=np.int64)
    beta = np.zeros(J); scores = np.zeros((J, M), order="F"); se = np.zeros((J, M), order="F")
    fail = np.zeros(J, dtype=np.int8); iters = np.zeros(J, dtype=np.int32)
    a = L.ProjArgs()
    a.Y = L.ptr(Yf); a.y_dtype = L.I32; a.I_full = Ik; a.J = J; a.ldY = Ik; a.ixs = L.ptr(ixs); a.Ik = Ik; a.M = M
    a.U = L.ptr(Uf); a.alpha = L.ptr(alpha); a.glm_tol = tol; a.glm_maxit = 100; a.device = -1; a.glm_method = method
    o = L.ProjOut(); o.beta = L.ptr(beta); o.scores = L.ptr(scores); o.se_scores = L.ptr(se)
    o.fail = L.ptr(fail); o.iters = L.ptr(iters)
    L.check(L.load().scgbm_project(C.byref(a), C.byref(o)))
    return dict(beta=beta, scores=scores, se=se, fail=fail, iters=iters, method=o.method_used, sweeps=o.tc_sweeps,
                stragglers=o.tc_stragglers)


@pytest.mark.parametrize("Ik,J,M,seed", [(700, 300, 6, 5), (1100, 520, 20, 6), (260, 200, 3, 7)])
def test_projection_tcgen05_matches_fastglm_restatement(Ik, J, M, seed):
    """K9 on the tensor cores (proj_tc.cu: lock-step Newton, gradient and Hessian from the fused pass's product rider)
    against the fastglm restatement, cell by cell.  Tolerances of this path (include/scgbm.h, scgbm_glm_method):
    coefficients 2e-6 (eta carries 22 bits, exp is ex2.approx), standard errors 1e-3 relative (fastglm reports the
    X^T W X of its LAST SOLVE, i.e. one step before the returned coefficients, so that number depends on the path
    by the size of the last step); the exact kernel holds 1e-9 for both (test_projection_glm_kernel_exact).
    A cell without counts, and the last percent of slow cells, take the exact kernel."""
    from oracle.scgbm_oracle import poisson_glm_irls
    U, alpha, Y = _glm_problem(Ik, J, M, seed)
    Y[:, 17] = 0                                            # no counts: handed to the exact kernel
    tc = _project(U, alpha, Y, 2)
    ex = _project(U, alpha, Y, 1)
    assert tc["method"] == 2 and ex["method"] == 1
    assert 1 <= tc["stragglers"] <= J // 20 + 2, tc              # the empty cell, and at most a few slow ones
    assert 3 <= tc["sweeps"] <= 30
    np.testing.assert_array_equal(tc["fail"], ex["fail"])
    ok = ex["fail"] == 0
    # against the exact kernel (itself pinned to the oracle at 1e-9) for every cell ...
    np.testing.assert_allclose(tc["beta"][ok], ex["beta"][ok], rtol=0, atol=2e-6)
    np.testing.assert_allclose(tc["scores"][ok], ex["scores"][ok], rtol=2e-6, atol=2e-6 * np.abs(ex["scores"][ok]).max())
    np.testing.assert_allclose(tc["se"][ok], ex["se"][ok], rtol=1e-3)
    # ... and against the oracle directly for a sample of them
    Xd = np.hstack([np.ones((Ik, 1)), U])
    for j in range(0, J, max(1, J // 12)):
        if not ok[j]:
            continue
        coef, sej, conv, nit = poisson_glm_irls(Xd, Y[:, j].astype(float), alpha)
        np.testing.assert_allclose(tc["beta"][j], coef[0], rtol=0, atol=2e-6)
        np.testing.assert_allclose(tc["scores"][j], coef[1:], rtol=2e-6, atol=2e-6 * np.abs(coef[1:]).max())
        np.testing.assert_allclose(tc["se"][j], sej[1:], rtol=1e-3)
        assert 1 <= int(tc["iters"][j]) <= nit + 2          # Newton solves from the null model vs fastglm's IRLS count


def test_projection_tcgen05_through_gbm_sc_subset():
    """gbm.sc(subset = idx) with the tensor-core projection against the exact one: same subsample fit, scores / beta /
    alpha within the path's tolerance; AUTO keeps the exact kernel below 4096 cells."""
    from scgbm_b200 import gbm_sc
    Y = _sim(260, 900, 4, seed=31, bmean=0.5)
    idx = np.arange(1, 301)
    ex = gbm_sc(Y, 4, subset=idx, glm_method="exact")
    tc = gbm_sc(Y, 4, subset=idx, glm_method="tcgen05")
    auto = gbm_sc(Y, 4, subset=idx)
    assert ex["_glm_method"] == "exact" and tc["_glm_method"] == "tcgen05" and auto["_glm_method"] == "exact"
    assert not tc["_fail"].any()
    np.testing.assert_allclose(tc["scores"], ex["scores"], rtol=2e-6, atol=2e-6 * np.abs(ex["scores"]).max())
    np.testing.assert_allclose(tc["se_scores"], ex["se_scores"], rtol=1e-3)
    np.testing.assert_allclose(tc["beta"], ex["beta"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(tc["alpha"], ex["alpha"], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(tc["U"], ex["U"])
