"""On-device generators (simData / simNull / simNullNB, R/simulate.R:24-65) and the CCI perturbation
sampler (8f-4, R/cci.R:69-71): seeded parity with the oracle's restatement of the stream (oracle/philox.py, pinned to
the published Philox known answers), and distribution tests of the counts against the oracle's own sim_data."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sim_data_factors_match_stream_restatement():
    from scgbm_b200 import sim_data
    from oracle import philox as P
    I, J, d = 300, 520, 4
    val = sim_data(I, J, d, alpha_mean=-0.5, beta_mean=0.25, K=3.0, seed=11)
    assert list(val)[:6] == ["Y", "V", "U", "D", "alpha", "beta"]                 # R/simulate.R:40-46
    U, V, D, alpha, beta = P.sim_data_factors(I, J, d, alpha_mean=-0.5, beta_mean=0.25, K=3.0, seed=11)
    np.testing.assert_allclose(val["D"], D, rtol=1e-14)
    np.testing.assert_allclose(val["alpha"], alpha, rtol=0, atol=1e-12)
    np.testing.assert_allclose(val["beta"], beta, rtol=0, atol=1e-12)
    np.testing.assert_allclose(val["U"], U, rtol=0, atol=1e-11)                    # Cholesky-QR on device == QR with diag(R) > 0
    np.testing.assert_allclose(val["V"], V * D[None, :], rtol=0, atol=1e-9)        # val$V = V diag(D)   (:42)
    assert np.all(val["U"][0] >= 0)                                                # :28-33
    assert np.abs(val["U"].T @ val["U"] - np.eye(d)).max() < 1e-12
    Vs = val["V"] / val["D"][None, :]
    assert np.abs(Vs.T @ Vs - np.eye(d)).max() < 1e-12
    assert val["Y"].shape == (I, J) and val["Y"].dtype == np.int32 and val["Y"].min() >= 0


def _moments(Y, mu, theta=None):
    """z-score of the total and the variance ratio against Poisson / NB(mu, theta)."""
    var = mu if theta is None else mu + mu * mu / theta
    z = (Y.sum() - mu.sum()) / np.sqrt(var.sum())
    ratio = ((Y - mu) ** 2).sum() / var.sum()
    return z, ratio


@pytest.mark.parametrize("theta", [0.0, 1.0, 3.0])
def test_sim_data_counts_follow_the_model(theta):
    """Y | Lambda ~ Poisson(exp(Lambda))  (R/simulate.R:37-38), or NB(mean exp(Lambda), theta) like MASS::rnegbin."""
    from scipy import stats
    from scgbm_b200 import sim_data
    from oracle.scgbm_oracle import sim_data as o_sim
    I, J, d = 400, 600, 3
    val = sim_data(I, J, d, alpha_mean=-1.0, K=2.0, seed=5, theta=theta)
    mu = np.exp(val["alpha"][:, None] + val["beta"][None, :] + val["U"] @ val["V"].T)
    Y = val["Y"].astype(np.float64)
    z, ratio = _moments(Y, mu, theta if theta > 0 else None)
    assert abs(z) < 4.5, z
    # the variance ratio is dominated by a few large means: bound it loosely, then test the bulk exactly
    assert 0.7 < ratio < 1.4, ratio
    # probability-integral transform on entries with a sizeable mean: randomised PIT of a discrete law is U(0, 1)
    rng = np.random.default_rng(0)
    m = (mu > 2) & (mu < 50)
    y, mm = Y[m][:20000], mu[m][:20000]
    if theta > 0:
        dist = stats.nbinom(theta, theta / (theta + mm))
    else:
        dist = stats.poisson(mm)
    pit = dist.cdf(y - 1) + rng.random(len(y)) * dist.pmf(y)
    assert stats.kstest(pit, "uniform").pvalue > 1e-3
    # against the oracle's own generator (numpy RNG, same recipe): same zero fraction and mean count
    ref = o_sim(I, J, d, alpha_mean=-1.0, seed=5, nb_theta=theta if theta > 0 else None)["Y"]
    assert abs((val["Y"] == 0).mean() - (ref == 0).mean()) < 0.03
    assert 0.5 < val["Y"].mean() / ref.mean() < 2.0          # heavy-tailed means: only the order of magnitude is stable


def test_shards_tile_the_matrix():
    """cells [j0, j1) of a J_total-cell matrix: bit-identical to the corresponding columns of the one-shot draw."""
    from scgbm_b200 import sim_data
    I, J, d = 150, 333, 3
    full = sim_data(I, J, d, seed=9, dtype=np.uint16)
    a = sim_data(I, 100, d, seed=9, dtype=np.uint16, j_offset=0, J_total=J)
    b = sim_data(I, J - 100, d, seed=9, dtype=np.uint16, j_offset=100, J_total=J)
    np.testing.assert_array_equal(np.hstack([a["Y"], b["Y"]]), full["Y"])
    np.testing.assert_array_equal(np.vstack([a["V"], b["V"]]), full["V"])
    np.testing.assert_array_equal(a["U"], full["U"]); np.testing.assert_array_equal(b["alpha"], full["alpha"])
    np.testing.assert_array_equal(np.concatenate([a["beta"], b["beta"]]), full["beta"])
    # a different seed is a different matrix
    assert (sim_data(I, J, d, seed=10, dtype=np.uint16)["Y"] != full["Y"]).mean() > 0.1


@pytest.mark.parametrize("theta", [None, 1.0, 0.5])
def test_sim_null_models(theta):
    """simNull / simNullNB (R/simulate.R:51-65): mu_i ~ U(1, 10) per gene, iid counts along the gene."""
    from scipy import stats
    from scgbm_b200 import sim_null, sim_null_nb
    from oracle import philox as P
    I, J = 600, 900
    if theta is None:
        Y, mu = sim_null(I, J, seed=4, return_mu=True)
    else:
        Y, mu = sim_null_nb(I, J, theta, seed=4, return_mu=True)
    np.testing.assert_allclose(mu, P.null_mus(I, seed=4), rtol=1e-13)
    assert mu.min() >= 1 and mu.max() <= 10 and stats.kstest((mu - 1) / 9, "uniform").pvalue > 1e-3
    Y = Y.astype(np.float64)
    M = np.repeat(mu[:, None], J, axis=1)
    z, ratio = _moments(Y, M, theta)
    assert abs(z) < 4.5 and abs(ratio - 1) < 0.05, (z, ratio)
    # gene by gene: chi-square goodness of fit of the row's counts against the model pmf
    ps = []
    for i in range(0, I, 40):
        dist = stats.poisson(mu[i]) if theta is None else stats.nbinom(theta, theta / (theta + mu[i]))
        kmax = int(dist.ppf(0.999)) + 1
        obs = np.bincount(np.minimum(Y[i].astype(int), kmax), minlength=kmax + 1).astype(float)
        exp = dist.pmf(np.arange(kmax + 1)); exp[-1] = 1 - exp[:-1].sum(); exp *= J
        keep = exp > 5
        o = np.append(obs[keep], obs[~keep].sum()); e = np.append(exp[keep], exp[~keep].sum())
        if e[-1] == 0:
            o, e = o[:-1], e[:-1]
        ps.append(stats.chisquare(o, e).pvalue)
    assert min(ps) > 1e-4 and stats.kstest(ps, "uniform").pvalue > 1e-3, ps


def test_cci_perturbation_sampling_matches_stream_restatement():
    """R/cci.R:69-71: V.new = V + rnorm(J*M, sd = se_V), `reps` times, as one batched device op."""
    from scipy import stats
    from scgbm_b200 import cci_perturb
    from oracle import philox as P
    rng = np.random.default_rng(2)
    J, M, reps = 700, 6, 9
    gbm = {"scores": np.asfortranarray(rng.normal(0, 5, (J, M))), "se_scores": np.asfortranarray(rng.uniform(0.05, 2.0, (J, M)))}
    out = cci_perturb(gbm, reps=reps, seed=21)
    assert out.shape == (J, M, reps)
    ref = P.cci_perturb(gbm["scores"], gbm["se_scores"], reps, seed=21)
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-12)
    # replicates are addressed by their index in the stream: batches fetched separately are the same draws
    np.testing.assert_array_equal(cci_perturb(gbm, reps=4, seed=21, rep_offset=3), out[:, :, 3:7])
    # the null draw of R/cci.R:57 is the perturbation alone
    e0 = cci_perturb(gbm, reps=reps, seed=21, null=True)
    np.testing.assert_allclose(e0, out - gbm["scores"][:, :, None], rtol=0, atol=1e-12)
    zs = (e0 / gbm["se_scores"][:, :, None]).ravel()
    assert abs(zs.mean()) < 0.02 and abs(zs.std() - 1) < 0.02 and stats.kstest(zs[:20000], "norm").pvalue > 1e-3
    assert abs(np.corrcoef(e0[:, :, 0].ravel(), e0[:, :, 1].ravel())[0, 1]) < 0.05      # replicates are independent
