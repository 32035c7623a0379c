"""Shared helpers for the parity tests (CUDA path vs oracle)."""
import numpy as np

# Tolerances stated by BASELINE.json's north_star:
TOL_OBJ_REL = 1e-6      # objective trace, relative
TOL_INTERCEPT = 1e-5    # alpha / beta, absolute
TOL_ANGLE = 1e-4        # principal angles of span(U), span(V), radians
KNIFE_EDGE = 1e-7       # below this oracle margin an accept/revert decision is not reproducible


def principal_angle(A, B):
    """largest principal angle between the column spans (radians)."""
    qa, _ = np.linalg.qr(A)
    qb, _ = np.linalg.qr(B)
    s = np.linalg.svd(qa.T @ qb, compute_uv=False)
    # 1 - s^2 loses precision near 0: use the sine through the complement
    r = qb - qa @ (qa.T @ qb)
    return float(np.arcsin(min(1.0, np.linalg.norm(r, 2))))


def first_knife_edge(ref):
    """index (0-based) of the first iteration whose accept/revert decision had an oracle
    margin below KNIFE_EDGE, or len if none."""
    LL = ref["_LL"]
    for i in range(2, len(LL)):
        if abs(LL[i] - LL[i - 1]) / abs(LL[i]) < KNIFE_EDGE:
            return i
    return len(LL)


def assert_fit_parity(ref, out, check_W=True, angle_tol=TOL_ANGLE, obj_tol=TOL_OBJ_REL, icpt_tol=TOL_INTERCEPT):
    k = first_knife_edge(ref)
    n = min(k, len(out["_LL"]), len(ref["_LL"]))
    assert n >= 1
    np.testing.assert_array_equal(ref["_accepted"][:n], out["_accepted"][:n])
    rel = np.abs(ref["_LL"][:n] - out["_LL"][:n]) / np.abs(ref["_LL"][:n])
    assert rel.max() < obj_tol, f"objective trace differs by {rel.max():.3e}"
    if k < len(ref["_LL"]):
        return "knife-edge"          # trace parity up to the first unreproducible decision only
    assert ref["_n_iter"] == out["_n_iter"]
    np.testing.assert_allclose(out["obj"], ref["obj"], rtol=obj_tol)
    assert np.abs(ref["alpha"] - out["alpha"]).max() < icpt_tol
    assert np.abs(ref["beta"] - out["beta"]).max() < icpt_tol
    assert principal_angle(ref["U"], out["U"]) < angle_tol
    assert principal_angle(ref["V"], out["V"]) < angle_tol
    np.testing.assert_allclose(out["D"], ref["D"], rtol=1e-5)
    # sign-aligned scores: the U[1,m] >= 0 convention fixes signs on both sides
    M = ref["U"].shape[1]
    gap = np.min(np.abs(np.diff(ref["D"])) / ref["D"][:-1]) if M > 1 else 1.0
    if gap > 1e-3:      # individual factors are only identified when the spectrum is separated
        for m in range(M):
            c = abs(np.dot(ref["U"][:, m], out["U"][:, m]))
            assert c > 1 - 1e-6
            assert np.dot(ref["scores"][:, m], out["scores"][:, m]) > 0
    assert np.all(out["U"][0, :] >= 0)
    np.testing.assert_allclose(out["scores"], out["V"] * out["D"][None, :], rtol=1e-12, atol=1e-300)
    if check_W and "W" in ref:
        np.testing.assert_allclose(out["W"], ref["W"], rtol=5e-4)
    return "full"
