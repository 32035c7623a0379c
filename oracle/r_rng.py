"""TEST INFRASTRUCTURE ONLY -- restatement of the slice of R's RNG needed to
regenerate the reference README demo input (README.md:17,24-27):

    set.seed(1126490984); Y <- matrix(rpois(I*J, lambda=1), nrow=I, ncol=J)

R itself is absent from this container, so this restates the published
algorithms of R's `src/main/RNG.c` (Mersenne-Twister seeding by `set.seed`)
and `src/nmath/rpois.c` (case B: table-lookup inversion for mu < 10) [ext].
It is self-validating: if anything here were wrong the oracle's first
objective would not print as the README's -240304.2 (tests/test_oracle_kat.py).

Nothing in the product imports this module.
"""
import numpy as np

_I2_32M1 = 2.328306437080797e-10  # R: fixup() bounds, 1/(2^32 - 1)


def r_set_seed_mt(seed: int) -> np.random.MT19937:
    """RNG_Init(MERSENNE_TWISTER, seed) [ext RNG.c]: 50 LCG scrambles, then 625
    LCG outputs fill i_seed (dummy[0]=mti, dummy[1..624]=mt); FixupSeeds forces
    mti = 624 so the first draw regenerates the state."""
    s = np.uint32(seed & 0xFFFFFFFF)
    a = np.uint32(69069)
    one = np.uint32(1)
    with np.errstate(over="ignore"):
        for _ in range(50):
            s = a * s + one
        dummy = np.empty(625, dtype=np.uint32)
        for j in range(625):
            s = a * s + one
            dummy[j] = s
    bg = np.random.MT19937()
    st = bg.state
    st["state"]["key"] = dummy[1:].copy()
    st["state"]["pos"] = 624
    bg.state = st
    return bg


def r_unif_rand(bg: np.random.MT19937, n: int) -> np.ndarray:
    """MT_genrand() * 2.3283064365386963e-10 then fixup() into (0,1) [ext RNG.c]."""
    raw = bg.random_raw(n).astype(np.float64)
    u = raw * 2.3283064365386963e-10
    u = np.where(u <= 0.0, 0.5 * _I2_32M1, u)
    u = np.where(1.0 - u <= 0.0, 1.0 - 0.5 * _I2_32M1, u)
    return u


def r_rpois_small(bg: np.random.MT19937, n: int, mu: float) -> np.ndarray:
    """rpois(n, mu) for 0 < mu < 10 [ext rpois.c, case B].  One uniform per
    draw; result is the smallest k with u <= pp[k], pp built by the running
    products p *= mu/k; q += p (the same floating-point order as R).  The
    u > pp[35] retry branch has probability < 1e-40 for mu = 1 and is guarded."""
    assert 0.0 < mu < 10.0
    p0 = np.exp(-mu)
    pp = np.empty(36)
    pp[0] = p0
    p = p0
    q = p0
    for k in range(1, 36):
        p *= mu / k
        q += p
        pp[k] = q
    u = r_unif_rand(bg, n)
    # u <= p0 -> 0 ; else first k>=1 with u <= pp[k]
    out = np.searchsorted(pp, u, side="left")
    if np.any(out > 35):
        raise RuntimeError("rpois retry branch hit; not restated")
    return out.astype(np.int32)


def readme_demo_Y(I: int = 500, J: int = 500, seed: int = 1126490984) -> np.ndarray:
    """README.md:17,24-27 -- column-major fill like R's matrix()."""
    bg = r_set_seed_mt(seed)
    y = r_rpois_small(bg, I * J, 1.0)
    return np.asfortranarray(y.reshape((I, J), order="F"))
