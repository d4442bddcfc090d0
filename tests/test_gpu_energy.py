"""Energy-matrix kernel (k_energy_rows, double precision) against the CPU oracle on identical float32-rounded
inputs.  Tolerance: 1e-5 relative (BASELINE.json north_star); the kernel is expected to do ~1e-12."""
import numpy as np
import pytest
from helpers import lj_setup, oracle_system, gpu_engine, KB

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def build(s, K, M, lambdas, temps, annihilate=False, c=6.0, a=1.0, b=1.0, offsets=None, rank=0, world=1):
    e = gpu_engine(1, K, M, s['N'], box=(s['L'],) * 3, r_cutoff=s['rc'], r_switch=s['rs'], use_switch=True,
                   annihilate_sterics=annihilate, softcore_c=c, softcore_a=a, softcore_b=b, rank=rank,
                   world_size=world)
    e.set_particles(s['sigma'], s['eps'], s['mass'], s['alch'])
    e.set_states(temps, lambdas, offsets)
    return e


@pytest.mark.parametrize('annihilate,c,a,b', [(False, 6.0, 1.0, 1.0), (True, 6.0, 1.0, 1.0), (False, 4.0, 2.0, 1.5)])
def test_energy_matrix_matches_oracle(annihilate, c, a, b):
    N, K, M = 512, 8, 64
    s = lj_setup(N=N, n_alch=10, seed=3)
    rng = np.random.default_rng(0)
    xs = np.stack([(s['x'] + rng.normal(scale=0.03, size=(N, 3))).astype(np.float32).astype(np.float64)
                   for _ in range(K)])
    lambdas = np.linspace(1.0, 0.0, M)
    temps = np.linspace(300.0, 330.0, M)
    offsets = np.linspace(-3.0, 2.0, M)
    e = build(s, K, M, lambdas, temps, annihilate, c, a, b, offsets)
    e.set_positions(xs)
    u = e.compute_energies()
    osys = oracle_system(s, annihilate=annihilate, c=c, a=a, b=b)
    ref = osys.energy_matrix(xs, lambdas, 1.0 / (KB * temps), offsets)
    err = np.abs(u - ref) / np.maximum(np.abs(ref), 1e-6 / KB / 300)
    assert err.max() < RTOL, err.max()
    # and an independent double loop through the single-configuration path (reference test_sampling.py:1668-1717)
    for k in (0, K - 1):
        row = osys.energy_row(xs[k], lambdas, 1.0 / (KB * temps), offsets)
        assert np.allclose(u[k], row, rtol=RTOL, atol=0)
    e.close()


def test_wrapped_positions_give_same_energies():
    """Periodic images: shifting atoms by box vectors must not change u."""
    N, K, M = 128, 2, 8
    s = lj_setup(N=N, n_alch=4, seed=5)
    lambdas = np.linspace(1, 0, M); temps = np.full(M, 300.0)
    e = build(s, K, M, lambdas, temps)
    xs = np.stack([s['x'], s['x']])
    shift = np.random.default_rng(1).integers(-2, 3, size=(N, 3)) * s['L']
    xs[1] += shift
    xs = xs.astype(np.float32).astype(np.float64)
    e.set_positions(xs)
    u = e.compute_energies()
    assert np.allclose(u[0], u[1], rtol=1e-4)
    # get_positions wraps into [0, L) like getState(enforcePeriodicBox=True)
    out = e.get_positions()
    assert out.min() >= 0 and out.max() < s['L'] * (1 + 1e-6)
    e.close()


def test_all_atoms_alchemical_and_none():
    N, K, M = 64, 2, 4
    for n_alch in (0, 64):
        s = lj_setup(N=N, n_alch=n_alch, seed=7)
        lambdas = np.array([1.0, 0.7, 0.3, 0.0]); temps = np.full(M, 300.0)
        for ann in (False, True):
            e = build(s, K, M, lambdas, temps, annihilate=ann)
            xs = np.stack([s['x'], s['x'][::-1].copy()])
            e.set_positions(xs)
            u = e.compute_energies()
            ref = oracle_system(s, annihilate=ann).energy_matrix(xs, lambdas, 1.0 / (KB * temps))
            assert np.allclose(u, ref, rtol=RTOL, atol=1e-9), (n_alch, ann)
            e.close()
