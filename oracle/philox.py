"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based random stream behind the on-device
generators (scgbm_b200/csrc/philox.cuh, simulate.cu) and, on top of it, of the reference's generator recipe
`simData` (/root/reference/R/simulate.R:24-48) and of CCI's perturbation step (/root/reference/R/cci.R:69-71).

Philox4x32-10 (Salmon et al., SC'11; the published algorithm, constants below) keyed by a 64-bit key with the
128-bit counter (idx_lo, idx_hi, block, 0x5C6B0000); the four 32-bit outputs of a block are consumed last to
first; a uniform takes two of them (53 bits), a normal is Box-Muller of two uniforms.  R's own RNG stream is NOT
what the device reproduces (the reference's draws are restated as distributions); what this module
pins is that the device draws exactly this documented stream, so seeded outputs are checkable bit-near."""
from __future__ import annotations

import math

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
KEY_STEP = 0x9E3779B97F4A7C15
FAM_U, FAM_V, FAM_ALPHA, FAM_BETA, FAM_MU, FAM_CCI = 1, 2, 3, 4, 5, 6
MASK = np.uint64(0xFFFFFFFF)


def fam_key(seed, fam):
    return (int(seed) + KEY_STEP * fam) & 0xFFFFFFFFFFFFFFFF


def philox_block(key, idx, block=0):
    """outputs (n, 4) uint64 (32-bit values) of Philox4x32-10 for counters idx (uint64 array)."""
    idx = np.asarray(idx, dtype=np.uint64)
    c0 = idx & MASK; c1 = idx >> np.uint64(32)
    c2 = np.full_like(idx, block); c3 = np.full_like(idx, 0x5C6B0000)
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF; k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1)


def _uniform_from(hi_word, lo_word):
    a = (hi_word >> np.uint64(5)).astype(np.float64); b = (lo_word >> np.uint64(6)).astype(np.float64)
    return (a * 67108864.0 + b + 0.5) * (1.0 / 9007199254740992.0)


def uniform(key, idx):
    o = philox_block(key, idx)
    return _uniform_from(o[:, 3], o[:, 2])


def normal(key, idx):
    o = philox_block(key, idx)
    u1 = _uniform_from(o[:, 3], o[:, 2]); u2 = _uniform_from(o[:, 1], o[:, 0])
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def normal_matrix(key, n, q, row_offset=0):
    """(n, q): element (r, c) uses counter (row_offset + r) * q + c."""
    r, c = np.meshgrid(np.arange(n, dtype=np.uint64), np.arange(q, dtype=np.uint64), indexing="ij")
    return normal(key, ((np.uint64(row_offset) + r) * np.uint64(q) + c).ravel()).reshape(n, q)


def haar(key, n, d):
    """Q of the QR factorisation with positive diag(R) of the (n, d) Gaussian block: Haar on the Stiefel manifold
    (the distribution of rstiefel::rustiefel, R/simulate.R:25)."""
    G = normal_matrix(key, n, d)
    Q, R = np.linalg.qr(G)
    return Q * np.sign(np.diag(R))[None, :]


def sim_data_factors(I, J, d, alpha_mean=0.0, beta_mean=0.0, K=2.0, seed=0):
    """U, V (unscaled), D, alpha, beta of R/simulate.R:25-36 from the device's stream."""
    U = haar(fam_key(seed, FAM_U), I, d); V = haar(fam_key(seed, FAM_V), J, d)
    for m in range(d):                                                  # :28-33
        if U[0, m] < 0:
            U[:, m] = -U[:, m]; V[:, m] = -V[:, m]
    s = math.sqrt(I) + math.sqrt(J)
    D = np.linspace(K * s, s, d) if d > 1 else np.array([K * s])        # :27
    alpha = alpha_mean + normal(fam_key(seed, FAM_ALPHA), np.arange(I, dtype=np.uint64))   # :35
    beta = beta_mean + normal(fam_key(seed, FAM_BETA), np.arange(J, dtype=np.uint64))      # :36
    return U, V, D, alpha, beta


def null_mus(I, seed=0):
    """mus <- runif(I, 1, 10)   (R/simulate.R:52,61)."""
    return 1.0 + 9.0 * uniform(fam_key(seed, FAM_MU), np.arange(I, dtype=np.uint64))


def cci_perturb(scores, se_scores, reps, seed=0, rep_offset=0):
    """V.new = V + rnorm(J*M, sd = se_V) for `reps` replicates (R/cci.R:69-71): (J, M, reps)."""
    se = np.asarray(se_scores, dtype=np.float64)
    J, M = se.shape
    n = J * M
    out = np.empty((J, M, reps))
    base = np.zeros((J, M)) if scores is None else np.asarray(scores, dtype=np.float64)
    e = np.arange(n, dtype=np.uint64)
    for r in range(reps):
        z = normal(fam_key(seed, FAM_CCI), np.uint64(rep_offset + r) * np.uint64(n) + e)
        out[:, :, r] = base + se * z.reshape((J, M), order="F")
    return out
