"""Pins the oracle's restatement of the generators' random stream (oracle/philox.py) against the published
known-answer vectors of Philox4x32-10 (Random123 kat_vectors) and checks the derived draws' moments."""
import numpy as np

from oracle import philox as P


def _block(ctr, key):
    c0, c1, c2, c3 = [np.array([x], dtype=np.uint64) for x in ctr]
    k0, k1 = key
    for _ in range(10):
        p0 = P.M0 * c0; p1 = P.M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & P.MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & P.MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + P.W0) & 0xFFFFFFFF; k1 = (k1 + P.W1) & 0xFFFFFFFF
    return [int(x[0]) for x in (c0, c1, c2, c3)]


def test_philox4x32_10_known_answers():
    assert _block((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _block((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _block((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # and the vectorised form used everywhere else is the same function (block 0, the stream's fixed 4th word)
    o = P.philox_block(0x299f31d0a4093822, np.array([0x85a308d3243f6a88], dtype=np.uint64))
    assert [int(v) for v in o[0]] == _block((0x243f6a88, 0x85a308d3, 0, 0x5C6B0000), (0xa4093822, 0x299f31d0))


def test_stream_moments_and_recipe():
    from scipy import stats
    z = P.normal(P.fam_key(7, P.FAM_ALPHA), np.arange(200000, dtype=np.uint64))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert stats.kstest(z[:20000], "norm").pvalue > 1e-3
    u = P.uniform(P.fam_key(7, P.FAM_MU), np.arange(20000, dtype=np.uint64))
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    U, V, D, alpha, beta = P.sim_data_factors(120, 90, 4, alpha_mean=-1.0, K=3.0, seed=3)   # R/simulate.R:24-36
    assert np.abs(U.T @ U - np.eye(4)).max() < 1e-12 and np.abs(V.T @ V - np.eye(4)).max() < 1e-12
    assert np.all(U[0] >= 0)
    s = np.sqrt(120) + np.sqrt(90)
    np.testing.assert_allclose(D, np.linspace(3 * s, s, 4))
    assert abs(alpha.mean() + 1.0) < 0.3 and abs(beta.mean()) < 0.3
    mus = P.null_mus(5000, seed=1)
    assert mus.min() >= 1.0 and mus.max() <= 10.0 and abs(mus.mean() - 5.5) < 0.1
