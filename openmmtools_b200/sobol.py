"""Sobol' low-discrepancy points (first three dimensions), written from the published algorithm
(Antonov-Saleev Gray-code recurrence with the Bratley-Fox initial direction numbers).

The reference seeds LennardJonesFluid with ``sobol.i4_sobol_generate(3, N, 1)``
(/root/reference/openmmtools/testsystems.py:277-284); tests/test_testsystems.py checks this module against points
captured from the reference's generator (tests/golden/sobol_golden.npz).
"""
import numpy as np

_BITS = 30
# dimension -> (primitive polynomial degree, coefficients a_1..a_{d-1}, initial m_1..m_d)
_INIT = [
    (0, (), ()),          # dim 1: van der Corput, all m = 1
    (1, (), (1,)),        # dim 2: x + 1
    (2, (1,), (1, 1)),    # dim 3: x^2 + x + 1
]


def _direction_numbers(dim):
    deg, coeffs, m0 = _INIT[dim]
    m = [1] * _BITS if deg == 0 else list(m0) + [0] * (_BITS - deg)
    for i in range(deg, _BITS):
        if deg == 0:
            break
        new = m[i - deg] ^ (m[i - deg] << deg)
        for k in range(1, deg):
            if coeffs[k - 1]:
                new ^= m[i - k] << k
        m[i] = new
    # v_j = m_j * 2^(BITS - j)
    return [m[j] << (_BITS - 1 - j) for j in range(_BITS)]


def sobol_generate(dim_num, n, skip=1):
    """Array of shape (dim_num, n) with the reference generator's numbering: ``skip=1`` starts at the origin
    (the reference's ``i4_sobol(dim, seed=1)`` returns the point of index 0)."""
    skip = max(int(skip) - 1, 0)
    if dim_num > 3:
        raise NotImplementedError('only the first three Sobol dimensions are provided')
    v = [_direction_numbers(d) for d in range(dim_num)]
    out = np.zeros((dim_num, n))
    x = [0] * dim_num
    scale = 1.0 / (1 << _BITS)
    # Gray-code recurrence: point k+1 = point k XOR v[position of the lowest zero bit of k]
    for k in range(skip + n):
        if k >= skip:
            for d in range(dim_num):
                out[d, k - skip] = x[d] * scale
        c = 0
        kk = k
        while kk & 1:
            kk >>= 1
            c += 1
        for d in range(dim_num):
            x[d] ^= v[d][c]
    return out


i4_sobol_generate = sobol_generate
