"""The oracle's Langevin splitting integrator and soft-core pair energy against vectors produced by the REAL
reference code (tests/golden/make_integrator_golden.py): openmmtools' ``LangevinIntegrator`` builds its step program
against a recording stand-in for ``openmm.CustomIntegrator`` and the program is interpreted in float64;
``AbsoluteAlchemicalFactory._get_sterics_energy_expressions`` supplies the Lepton energy expression.  This pins the
integrator algebra (sub-step formulas, a = exp(-gamma h), b = sqrt(1 - exp(-2 gamma h)) with h = dt / n_O,
sigma = sqrt(kT/m), the dt / n_V and dt / n_R fractions of a splitting string) and the soft-core formula with its
mixing rules to the reference's own text; the force field evaluation itself remains OpenMM's."""
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from helpers import oracle_system, KB
from make_integrator_golden import CASES, STERICS_GRID, case_inputs   # pure-python tables and input builders

G = np.load(os.path.join(HERE, 'golden', 'integrator_golden.npz'))


def test_reference_program_is_the_documented_one():
    prog = [str(p) for p in G['program_VRORV']]
    assert 'perdof|v|v + (dt / 2) * f / m' in prog and 'perdof|x|x + ((dt / 2) * v)' in prog
    assert 'perdof|v|(a * v) + (b * sigma * gaussian)' in prog and 'perdof|sigma|sqrt(kT/m)' in prog
    assert str(G['sterics_expression'][1]).startswith('U_sterics;U_sterics = ((lambda_sterics)^softcore_a)*4*epsilon*x*(x-1.0)')


@pytest.mark.parametrize('idx', range(len(CASES)))
def test_oracle_langevin_matches_the_interpreted_reference_program(idx):
    splitting, n_steps, dt, gamma, T, lam = CASES[idx]
    s, v0, noise = case_inputs(idx)
    kT, a, b = G['case%d_globals' % idx]
    h = dt / splitting.split().count('O')
    assert abs(kT - KB * T) < 1e-12 * kT
    assert abs(a - np.exp(-gamma * h)) < 1e-15 and abs(b - np.sqrt(1.0 - np.exp(-2.0 * gamma * h))) < 1e-15
    x = s['x'].copy(); v = v0.copy()
    oracle_system(s).langevin(x, v, noise, lam, KB * T, dt, gamma, n_steps, splitting.replace(' ', ''))
    assert np.abs(x - G['case%d_x' % idx]).max() < 1e-12
    assert np.abs(v - G['case%d_v' % idx]).max() < 1e-11


def test_oracle_softcore_pair_energy_matches_the_reference_expression():
    from oracle import oracle
    g = STERICS_GRID
    U = G['sterics_U']
    worst = 0.0
    for p, (s1, s2, e1, e2, alpha, a, b, c) in enumerate(g['params']):
        osys = oracle.LJSystem(np.array([s1, s2]), np.array([e1, e2]), np.array([39.9, 39.9]), np.array([1, 0], np.uint8),
                               (10.0, 10.0, 10.0), 4.0, 3.9, use_switch=False, alpha=alpha, a=a, b=b, c=c)
        for l, lam in enumerate(g['lam']):
            for q, r in enumerate(g['r']):
                xyz = np.array([[1.0, 1.0, 1.0], [1.0 + r, 1.0, 1.0]])
                e = osys.energy(xyz, lam)[0]
                ref = U[p, l, q]
                worst = max(worst, abs(e - ref) / max(1.0, abs(ref)))
    assert worst < 1e-12, worst
