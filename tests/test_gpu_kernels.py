"""Kernel-level parity (through the C ABI's diagnostic entry points) against numpy."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from scgbm_b200 import _lib as L
    return L.load()


@pytest.mark.parametrize("n", [1, 2, 5, 31, 32, 44, 63, 96, 97, 128, 180, 129, 160, 311])
def test_eigh_matches_lapack(lib, n):
    from scgbm_b200 import _lib as L
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    A = A @ A.T + (np.diag(rng.standard_normal(n)) if n % 2 else 0)
    A = (A + A.T) / 2
    Af = np.asfortranarray(A.copy())
    w = np.zeros(n)
    L.check(lib.scgbm_dbg_eigh(L.ptr(Af), L.ptr(w), n, -1))
    wr = np.linalg.eigvalsh(A)[::-1]
    scale = np.abs(wr).max()
    assert np.abs(w - wr).max() < 1e-12 * scale
    assert np.abs(A @ Af - Af * w).max() < 1e-12 * scale
    assert np.abs(Af.T @ Af - np.eye(n)).max() < 1e-12


def test_eigh_rank_deficient(lib):
    from scgbm_b200 import _lib as L
    rng = np.random.default_rng(0)
    B = rng.standard_normal((40, 7))
    A = B @ B.T
    Af = np.asfortranarray(A.copy()); w = np.zeros(40)
    L.check(lib.scgbm_dbg_eigh(L.ptr(Af), L.ptr(w), 40, -1))
    assert np.abs(w[7:]).max() < 1e-10 * w[0]
    assert np.abs(A @ Af - Af * w).max() < 1e-11 * w[0]


def _tsvd(lib, R32, M, terms=(), c=1.0, tol=1e-6, extra=0, depth=0, V0=None, seed=0, variant=0):
    from scgbm_b200 import _lib as L
    I, J = R32.shape if R32 is not None else (terms[0][0].shape[0], terms[0][2].shape[0])
    a = L.TsvdArgs()
    a.I = I; a.J = J; a.R = L.ptr(R32); a.c = c; a.nterms = len(terms); a.M = M
    keep = []
    for t, (A, d, B, coef) in enumerate(terms):
        A = np.asfortranarray(A, dtype=np.float64); B = np.asfortranarray(B, dtype=np.float64)
        d = np.ascontiguousarray(d, dtype=np.float64)
        keep += [A, B, d]
        a.A[t] = A.ctypes.data; a.d[t] = d.ctypes.data; a.B[t] = B.ctypes.data
        a.Mt[t] = A.shape[1]; a.coef[t] = coef
    a.svd_extra = extra; a.svd_depth = depth; a.svd_tol = tol; a.seed = seed; a.device = -1
    a.kernel_variant = variant
    if V0 is not None:
        V0 = np.asfortranarray(V0, dtype=np.float64); a.V0 = L.ptr(V0)
    U = np.zeros((I, M), order="F"); D = np.zeros(M); V = np.zeros((J, M), order="F")
    o = L.TsvdOut(); o.U = L.ptr(U); o.D = L.ptr(D); o.V = L.ptr(V)
    L.check(lib.scgbm_dbg_tsvd(C.byref(a), C.byref(o)))
    return U, D, V, o.passes, o.est


@pytest.mark.parametrize("shape,M", [((300, 800), 5), ((700, 1900), 10), ((1025, 333), 20), ((64, 5000), 8)])
def test_tsvd_dense_matches_lapack(lib, shape, M):
    from tests.util import principal_angle
    I, J = shape
    rng = np.random.default_rng(I + J)
    r = M + 3
    G = (rng.standard_normal((I, r)) * np.linspace(3.0, 1.0, r)) @ rng.standard_normal((J, r)).T / np.sqrt(r) \
        + 0.3 * rng.standard_normal((I, J))
    R32 = np.asfortranarray(G.astype(np.float32))
    U, D, V, passes, est = _tsvd(lib, R32, M, tol=1e-7)
    ue, de, vte = np.linalg.svd(R32.astype(np.float64), full_matrices=False)
    assert np.abs(D - de[:M]).max() < 1e-6 * de[0]
    assert principal_angle(ue[:, :M], U) < 1e-5
    assert principal_angle(vte[:M].T, V) < 1e-5
    assert np.abs(U.T @ U - np.eye(M)).max() < 1e-9
    Xe = (ue[:, :M] * de[:M]) @ vte[:M]
    Xa = (U * D) @ V.T
    assert np.linalg.norm(Xa - Xe) < 1e-5 * np.linalg.norm(Xe)


def test_tsvd_lowrank_terms_plus_residual(lib):
    """G = 1.25 A1 d1 B1' - 0.25 A2 d2 B2' + c R : the operator of R/fit.R:319-323."""
    from tests.util import principal_angle
    rng = np.random.default_rng(5)
    I, J, M = 400, 900, 6
    A1, _ = np.linalg.qr(rng.standard_normal((I, M))); B1, _ = np.linalg.qr(rng.standard_normal((J, M)))
    A2, _ = np.linalg.qr(A1 + 0.1 * rng.standard_normal((I, M))); B2, _ = np.linalg.qr(B1 + 0.1 * rng.standard_normal((J, M)))
    d1 = np.linspace(50, 20, M); d2 = np.linspace(48, 21, M)
    R = rng.standard_normal((I, J)).astype(np.float32); c = 0.07
    G = 1.25 * (A1 * d1) @ B1.T - 0.25 * (A2 * d2) @ B2.T + c * R.astype(np.float64)
    U, D, V, passes, est = _tsvd(lib, np.asfortranarray(R), M, terms=[(A1, d1, B1, 1.25), (A2, d2, B2, -0.25)], c=c, tol=1e-7)
    ue, de, vte = np.linalg.svd(G, full_matrices=False)
    assert np.abs(D - de[:M]).max() < 1e-7 * de[0]
    assert principal_angle(ue[:, :M], U) < 1e-6
    assert principal_angle(vte[:M].T, V) < 1e-6


def test_tsvd_warm_start_needs_fewer_passes(lib):
    rng = np.random.default_rng(9)
    I, J, M = 500, 1200, 8
    G = rng.standard_normal((I, 10)) @ rng.standard_normal((J, 10)).T / 3 + rng.standard_normal((I, J))
    R32 = np.asfortranarray(G.astype(np.float32))
    U, D, V, p_cold, _ = _tsvd(lib, R32, M, tol=1e-5, extra=12)
    ue, de, vte = np.linalg.svd(R32.astype(np.float64), full_matrices=False)
    V0 = vte[: M + 12].T + 1e-4 * rng.standard_normal((J, M + 12))
    U2, D2, V2, p_warm, _ = _tsvd(lib, R32, M, tol=1e-5, extra=12, V0=V0)
    assert p_warm <= 6 and p_warm < p_cold
    assert np.abs(D2 - de[:M]).max() < 1e-6 * de[0]


def test_tsvd_exactly_low_rank_and_tiny(lib):
    """rank(G) < block size, and a matrix smaller than the Krylov block."""
    rng = np.random.default_rng(2)
    G = (rng.standard_normal((90, 3)) @ rng.standard_normal((70, 3)).T).astype(np.float32)
    U, D, V, _, _ = _tsvd(lib, np.asfortranarray(G), 3, tol=1e-8)
    de = np.linalg.svd(G.astype(np.float64), compute_uv=False)
    assert np.abs(D - de[:3]).max() < 1e-6 * de[0]
    G = rng.standard_normal((9, 13)).astype(np.float32)
    U, D, V, _, _ = _tsvd(lib, np.asfortranarray(G), 4, tol=1e-8)
    de = np.linalg.svd(G.astype(np.float64), compute_uv=False)
    assert np.abs(D - de[:4]).max() < 1e-6 * de[0]


def _eval(lib, Y, Xu, Xv, beta, infer_beta=False, clamp=False, rowscale=None, colscale=None, storage=0):
    from scgbm_b200 import _lib as L
    from scgbm_b200.api import _as_counts
    Yh, dt = _as_counts(Y)
    I, J = Yh.shape
    a = L.EvalArgs()
    a.Y = L.ptr(Yh); a.y_dtype = dt; a.I = I; a.J = J; a.ldY = I; a.y_storage = storage
    Xu = np.asfortranarray(Xu, dtype=np.float64); Xv = np.asfortranarray(Xv, dtype=np.float64)
    a.Xu = L.ptr(Xu); a.Xv = L.ptr(Xv); a.Mx = Xu.shape[1]; a.clamp8 = int(clamp)
    rs = None if rowscale is None else np.ascontiguousarray(rowscale, dtype=np.float64)
    cs = None if colscale is None else np.ascontiguousarray(colscale, dtype=np.float64)
    a.rowscale = L.ptr(rs); a.colscale = L.ptr(cs)
    beta = np.ascontiguousarray(beta, dtype=np.float64); a.beta_in = L.ptr(beta)
    a.infer_beta = int(infer_beta); a.device = -1
    alpha_o = np.zeros(I); beta_o = np.zeros(J); resid = np.zeros((I, J), order="F")
    o = L.EvalOut(); o.alpha = L.ptr(alpha_o); o.beta = L.ptr(beta_o); o.resid = L.ptr(resid)
    L.check(lib.scgbm_dbg_eval(C.byref(a), C.byref(o)))
    return alpha_o, beta_o, resid, o


def _eval_ref(Y, X, beta, infer_beta):
    """R/fit.R:127-151 in float64."""
    Y = Y.astype(np.float64)
    alpha = np.log(Y.sum(1)) - np.log(np.exp(X + beta[None, :]).sum(1))
    am = alpha.mean(); beta = beta + am; alpha = alpha - am
    if infer_beta:
        beta = np.log(Y.sum(0)) - np.log(np.exp(alpha[:, None] + X).sum(0))
    W = np.exp(alpha[:, None] + X + beta[None, :])
    mY = Y.max()
    W[W > mY] = mY
    nz = Y != 0
    LL = (Y[nz] * np.log(W[nz])).sum() - W.sum()
    return alpha, beta, W, LL


@pytest.mark.parametrize("I,J,M,storage,dtype,infer", [
    (500, 700, 10, 0, np.int32, False), (129, 65, 3, 1, np.int32, True), (1000, 333, 20, 2, np.float64, False),
    (77, 2049, 7, 3, np.int32, True), (640, 640, 33, 0, np.uint16, False)])
def test_fused_pass_matches_float64(lib, I, J, M, storage, dtype, infer):
    rng = np.random.default_rng(I * 7 + J)
    Xu = rng.standard_normal((I, M)) * 0.4; Xv = rng.standard_normal((J, M)) * 0.4
    X = Xu @ Xv.T
    Y = rng.poisson(np.exp(0.3 * rng.standard_normal((I, 1)) + 0.3 * rng.standard_normal((1, J)) + X))
    Y[0, :] += 1; Y[:, 0] += 1
    Y = np.minimum(Y, 250)
    beta = np.log(Y.sum(0))
    a_o, b_o, resid, o = _eval(lib, Y.astype(dtype), Xu, Xv, beta, infer_beta=infer, storage=storage)
    a_r, b_r, W, LL = _eval_ref(Y, X, beta, infer)
    assert np.abs(a_o - a_r).max() < 5e-6
    assert np.abs(b_o - b_r).max() < 5e-6
    nz = Y != 0
    syl = (Y[nz] * np.log(W[nz])).sum()
    assert abs(o.sum_ylogw - syl) < 1e-7 * max(abs(syl), W.sum())
    assert abs(o.LL - LL) < 1e-7 * (abs(syl) + W.sum())
    assert abs(o.max_w - W.max()) < 1e-5 * W.max()
    assert abs(o.sum_w - W.sum()) < 1e-7 * W.sum()
    assert o.max_Y == Y.max()
    np.testing.assert_allclose(resid, Y - W, rtol=0, atol=2e-5 * max(1.0, W.max()))


def test_fused_pass_clamped_x0_form(lib):
    """X0 = clamp(s_i t_j (u d v')_ij, +-8)  (R/fit.R:116-120) evaluated from scaled factors."""
    rng = np.random.default_rng(3)
    I, J, M = 300, 500, 5
    Xu = rng.standard_normal((I, M)) * 2.0; Xv = rng.standard_normal((J, M)) * 2.0
    s = np.exp(0.3 * rng.standard_normal(I)); t = np.exp(0.3 * rng.standard_normal(J))
    X = np.clip((s[:, None] * t[None, :]) * (Xu @ Xv.T), -8, 8)
    assert (np.abs(X) == 8).mean() > 0.01
    Y = rng.poisson(2.0, size=(I, J)); Y[0, :] += 1; Y[:, 0] += 1
    beta = np.log(Y.sum(0))
    a_o, b_o, resid, o = _eval(lib, Y.astype(np.int32), Xu, Xv, beta, clamp=True, rowscale=s, colscale=t)
    a_r, b_r, W, LL = _eval_ref(Y, X, beta, False)
    assert np.abs(a_o - a_r).max() < 1e-5
    assert abs(o.LL - LL) < 1e-6 * abs(LL)
    assert abs(o.max_w - W.max()) <= 1e-6 * W.max()      # clipping at max.Y is active here


@pytest.mark.parametrize("storage", [2, 3])          # u16 -> tcgen05 pass + underflow guard; f32 -> CUDA-core pass
@pytest.mark.parametrize("x00,expect_inf", [(-800.0, True), (-720.0, False)])
def test_exp_underflow_gives_minus_inf_like_the_reference(lib, storage, x00, expect_inf):
    """R/fit.R:136,151-157: below eta = -745 R's exp() underflows to 0, log(0) = -Inf for a non-zero count and the
    objective is -Inf (the loop then reverts and halves lr).  The tensor-core pass sums Y * eta and cannot see
    that; its guard (eta_tc.cu ll_const_kernel: min alpha + min beta - |x| bound < -700) hands such an iteration
    to the CUDA-core pass, which restates the underflow per entry.  x = -720 trips the guard but not the
    underflow: the objective must then be the finite float64 value."""
    rng = np.random.default_rng(3)
    I, J = 96, 130
    Y = rng.poisson(2.0, size=(I, J)).astype(np.int32)
    Y[0, 0] = 3
    Xu = np.zeros((I, 1)); Xv = np.zeros((J, 1))
    Xu[0, 0] = x00; Xv[0, 0] = 1.0                    # X[0,0] = x00, every other entry 0
    beta = np.log(Y.sum(0))
    a_o, b_o, resid, o = _eval(lib, Y, Xu, Xv, beta, storage=storage)
    with np.errstate(divide="ignore"):
        a_r, b_r, W, LL = _eval_ref(Y, Xu @ Xv.T, beta, False)
    assert np.isneginf(LL) == expect_inf
    if expect_inf:
        assert np.isneginf(o.LL) and np.isneginf(o.sum_ylogw)
    else:
        assert abs(o.LL - LL) < 1e-6 * abs(LL)
    assert np.abs(a_o - a_r).max() < 5e-6


def test_nan_inf_objective_branch_matches_reference(monkeypatch):
    """Control flow of R/fit.R:152-157 and quirks Q1/Q2, pinned by replacing the objective of chosen iterations on
    BOTH sides (oracle `_ll_inject`, library SCGBM_DBG_LL): -Inf reverts X <- Xt and halves lr (below 1, which
    the LL-decrease rule never does), consumes the loop index, and leaves no entry in `obj`; a NaN objective
    makes the NEXT iteration's `if (LL[i] < LL[i-1])` an R error ("missing value where TRUE/FALSE needed")."""
    from scgbm_b200 import gbm_sc, RError
    from scgbm_b200._lib import ScgbmError
    from oracle.scgbm_oracle import gbm_sc as o_gbm_sc, sim_data, RError as ORError
    from tests.util import assert_trace_parity
    Y = sim_data(180, 260, 4, seed=17)["Y"]
    # (-Inf only, and never twice in a row: after a +Inf, or after two rejected trips in a row, the next trip
    # re-evaluates the very same X, LL[i] == LL[i-1] up to rounding, and `LL[i] < LL[i-1]` is a coin flip on
    # both sides -- a knife edge by construction, not a property of the branch under test)
    monkeypatch.setenv("SCGBM_DBG_LL", "2:-inf,6:-inf,10:-inf")
    ref = o_gbm_sc(Y, 4, max_iter=14, _ll_inject={2: -np.inf, 6: -np.inf, 10: -np.inf})
    out = gbm_sc(Y, 4, max_iter=14, diagnostics=True)
    assert not out["_accepted"][1] and not out["_accepted"][5] and not out["_accepted"][9]
    np.testing.assert_array_equal(ref["_accepted"], out["_accepted"])
    fin = np.isfinite(ref["_LL"])
    np.testing.assert_array_equal(np.isfinite(out["_LL"]), fin)
    rel = np.abs(ref["_LL"][fin] - out["_LL"][fin]) / np.abs(ref["_LL"][fin])
    assert rel.max() < 1e-6                          # the halved learning rates (0.5, 0.26, ...) reach the later steps
    assert len(out["obj"]) == len(ref["obj"]) == int(ref["_accepted"].sum())
    np.testing.assert_allclose(out["alpha"], ref["alpha"], atol=1e-5)
    # Q1: NaN at iteration i -> iteration i+1 compares against NA -> R stops
    monkeypatch.setenv("SCGBM_DBG_LL", "4:nan")
    with pytest.raises(ORError, match="missing value"):
        o_gbm_sc(Y, 4, max_iter=14, _ll_inject={4: np.nan})
    with pytest.raises(ScgbmError, match="missing value") as ei:
        gbm_sc(Y, 4, max_iter=14)
    assert ei.value.status == 3
