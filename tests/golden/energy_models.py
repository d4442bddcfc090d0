"""Deterministic synthetic reduced-energy matrices shared by the golden-vector generator and the tests.

Only +,-,*,/ on IEEE doubles and integer hashing are used, so the matrices are bit-identical on every
machine/numpy version and need not be stored in the fixtures.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)

def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))

def uniforms(n, key):
    """n doubles in [0,1) from a counter-based hash; exact (k / 2**53)."""
    with np.errstate(over='ignore'):
        ctr = np.arange(n, dtype=np.uint64) + (np.uint64(key) << np.uint64(32))
        z = _splitmix64(ctr)
    return (z >> np.uint64(11)).astype(np.float64) / 9007199254740992.0

def pseudo_normal(n, key):
    """Irwin-Hall(12) - 6: mean 0, variance 1, no transcendental functions."""
    u = uniforms(12 * n, key).reshape(n, 12)
    return u.sum(axis=1) - 6.0

def energies(model, K, key):
    if model == 'zeros':
        return np.zeros((K, K))
    if model == 'normal':
        return (2.0 * pseudo_normal(K * K, key)).reshape(K, K)
    if model == 'ladder':      # harmonic-overlap ladder: low acceptance, neighbour dominated
        mu = 0.5 * np.arange(K, dtype=np.float64)
        x = mu + pseudo_normal(K, key)
        return 0.5 * (x[:, None] - mu[None, :]) ** 2
    if model == 'flat':        # fine alchemical ladder: high acceptance
        lam = 1.0 - np.arange(K, dtype=np.float64) / max(K - 1, 1)
        a = 2.0 * pseudo_normal(K, key)
        return -600.0 + a[:, None] * lam[None, :]
    raise ValueError(model)


def ladder_energies(it, K, M, key):
    """u[k, l] of iteration `it` for the SAMS / online-analysis goldens: harmonic-ish ladder plus iteration
    dependent noise (exact arithmetic only)."""
    import numpy as np
    x = pseudo_normal(K, key * 1000 + it)
    mu = 0.7 * np.arange(M, dtype=np.float64)
    return 0.5 * (3.0 * x[:, None] - mu[None, :]) ** 2 * 0.1 + 0.05 * mu[None, :]


def ladder_states(it, K, M, key):
    """A deterministic replica -> state map of iteration `it` (not a permutation: the online update does not need one)."""
    import numpy as np
    return (np.floor(np.abs(pseudo_normal(K, key * 77 + it)) * 1e6).astype(np.int64) + it) % M
