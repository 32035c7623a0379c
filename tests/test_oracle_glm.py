"""CPU: the lock-step Newton iteration of the tensor-core projection path (proj_tc.cu), restated in float64
(oracle.poisson_glm_newton), against the fastglm restatement (oracle.poisson_glm_irls) -- the ALGORITHM's claims,
checked without a GPU: same maximum-likelihood coefficients, standard errors that differ by the size of the last
step, fewer solves from the null-model start on sparse counts."""
import numpy as np
import pytest

from oracle.scgbm_oracle import poisson_glm_irls, poisson_glm_newton


def _problem(Ik, M, seed, scale, off_mean=-1.0):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((Ik, M)))
    X = np.hstack([np.ones((Ik, 1)), U])
    off = rng.normal(off_mean, 0.8, Ik)
    return X, off, rng


# the last two cases are the benchmark's regime: ~65-90 % zeros, moderate scores
@pytest.mark.parametrize("Ik,M,seed,scale,off_mean,sparse", [(400, 5, 0, 3.0, -1.0, False), (900, 12, 1, 6.0, -1.0, False),
                                                             (300, 3, 2, 10.0, -1.0, False), (2000, 10, 3, 2.0, -2.5, True),
                                                             (2000, 20, 4, 1.0, -2.5, True)])
def test_newton_from_the_null_model_finds_fastglms_coefficients(Ik, M, seed, scale, off_mean, sparse):
    X, off, rng = _problem(Ik, M, seed, scale, off_mean)
    n_irls, n_newton = [], []
    for _ in range(12):
        theta = np.concatenate([[rng.normal(0.3, 0.6)], rng.normal(0, scale, M)])
        y = rng.poisson(np.exp(off + X @ theta)).astype(float)
        if y.sum() == 0:
            continue
        c1, s1, ok1, n1 = poisson_glm_irls(X, y, off)
        c2, s2, ok2, n2 = poisson_glm_newton(X, y, off)
        assert ok1 and ok2
        np.testing.assert_allclose(c2, c1, rtol=0, atol=2e-7 * max(1.0, np.abs(c1).max()))   # the same MLE
        np.testing.assert_allclose(s2, s1, rtol=1e-3)      # X'WX one step before the returned coefficients, on either path
        assert n2 <= n1 + 2
        n_irls.append(n1); n_newton.append(n2)
    if sparse:      # the null model is the better start for sparse counts: ~4 solves against fastglm's ~6
        assert np.mean(n_newton) <= np.mean(n_irls) - 1.0


def test_newton_rejects_a_cell_without_counts():
    X, off, _ = _problem(200, 3, 5, 1.0)
    with pytest.raises(FloatingPointError):
        poisson_glm_newton(X, np.zeros(200), off)
