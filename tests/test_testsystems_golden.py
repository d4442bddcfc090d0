"""openmmtools_b200.testsystems against the REAL reference classes (tests/golden/make_testsystems_golden.py:
``LennardJonesFluid`` / ``HarmonicOscillator`` / ``subrandom_particle_positions`` lifted by AST and run on recording
stand-ins for the OpenMM classes, with the reference's own sobol.py)."""
import json
import os
import numpy as np
import pytest
from openmmtools_b200 import testsystems, unit

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'testsystems_golden.npz'))

CASES = {
    'lj512': dict(nparticles=512),
    'lj100_dense': dict(nparticles=100, reduced_density=0.3, switch_width=2.0 * unit.angstroms, cutoff=9.0 * unit.angstroms),
    'lj64_noswitch': dict(nparticles=64, switch_width=None, dispersion_correction=False),
}


@pytest.mark.parametrize('tag', sorted(CASES))
def test_lennard_jones_fluid_matches_the_reference_class(tag):
    fluid = testsystems.LennardJonesFluid(**CASES[tag])
    s = fluid.system
    method, cutoff, use_switch, switch_distance, dispersion = G[tag + '_nb']
    assert np.allclose(s.masses, G[tag + '_mass'], rtol=1e-12) and np.allclose(s.sigma, G[tag + '_sigma'], rtol=1e-12)
    assert np.allclose(s.epsilon, G[tag + '_epsilon'], rtol=1e-12)
    assert np.all(G[tag + '_charge'] == 0.0)
    assert np.allclose(np.asarray(s.box_vectors), G[tag + '_box'], rtol=1e-12)
    assert method == 2 and s.box_vectors is not None                       # NonbondedForce.CutoffPeriodic
    assert abs(s.cutoff - cutoff) < 1e-12
    assert bool(s.use_switching_function) == bool(use_switch)
    if use_switch:
        assert abs(s.switching_distance - switch_distance) < 1e-12
    assert bool(s.use_dispersion_correction) == bool(dispersion)
    x = np.asarray(fluid.positions.value_in_unit(unit.nanometer), np.float64)
    assert x.shape == G[tag + '_positions'].shape
    assert np.abs(x - G[tag + '_positions']).max() < 1e-6 * G[tag + '_box'][0, 0]   # float32 positions (testsystems.py:267)


def test_harmonic_oscillator_matches_the_reference_class():
    ho = testsystems.HarmonicOscillator()
    s = ho.system
    g = json.loads(str(G['ho_globals']))
    assert str(G['ho_expression']).startswith('(K/2.0) * ((x-x0)^2 + y^2 + z^2) + U0;')
    assert abs(s.ho_K - g['testsystems_HarmonicOscillator_K']) < 1e-9 * s.ho_K
    assert s.ho_x0[0] == g['testsystems_HarmonicOscillator_x0'] == 0.0 and s.ho_U0 == g['testsystems_HarmonicOscillator_U0'] == 0.0
    assert np.allclose(s.masses, G['ho_mass'], rtol=1e-12)
    assert np.array_equal(np.asarray(ho.positions.value_in_unit(unit.nanometer)), G['ho_positions'])
