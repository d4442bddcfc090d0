"""The oracle's alchemical energy decomposition against the REAL reference factory (tests/golden/make_alchemy_golden.py):
``AbsoluteAlchemicalFactory._alchemically_modify_NonbondedForce`` (alchemy.py:1539-2038, lifted by AST) builds its
forces on recording stand-ins for the OpenMM force classes; they are evaluated with numpy (switching function off) and
described structurally.  Pins: which pairs go to which force (E-E in the NonbondedForce with the alchemical epsilons
zeroed, E-A soft core controlled by lambda_sterics, A-A soft core with lambda fixed to 1 unless annihilating), the
soft-core parameters handed over, and the long-range-correction flags."""
import json
import os
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from helpers import lj_setup
from make_alchemy_golden import CONFIGS, LAMBDAS, N, N_ALCH

G = np.load(os.path.join(HERE, 'golden', 'alchemy_golden.npz'))


@pytest.mark.parametrize('c', range(len(CONFIGS)))
def test_oracle_energy_equals_the_reference_force_decomposition(c):
    from oracle import oracle
    annihilate, disable_lrc, (alpha, a, b, cc) = CONFIGS[c]
    s = lj_setup(N=N, n_alch=N_ALCH, reduced_density=0.4, seed=77)
    assert np.array_equal(s['x'], G['x']) and s['L'] == float(G['L'])
    osys = oracle.LJSystem(s['sigma'], s['eps'], s['mass'], s['alch'], (s['L'],) * 3, s['rc'], s['rs'], use_switch=False,
                           alpha=alpha, a=a, b=b, c=cc, annihilate_sterics=annihilate)
    for lam, ref in zip(LAMBDAS, G['config%d_U' % c]):
        e = osys.energy(s['x'], lam)[0]
        assert abs(e - ref) < 1e-10 * max(1.0, abs(ref)), (c, lam, e, ref)


@pytest.mark.parametrize('c', range(len(CONFIGS)))
def test_reference_force_table_structure(c):
    annihilate, disable_lrc, (alpha, a, b, cc) = CONFIGS[c]
    desc = json.loads(str(G['config%d_forces' % c]))
    sterics = [f for f in desc['lambda_sterics'] if f['type'] == 'CustomNonbondedForce']
    na, aa = sterics
    assert na['groups'][0][:2] == [N - N_ALCH, N_ALCH] and aa['groups'][0][:2] == [N_ALCH, N_ALCH]
    assert 'lambda_sterics' in na['globals']                        # E-A pairs are lambda controlled
    assert ('lambda_sterics' in aa['globals']) == annihilate        # A-A pairs only when annihilating (alchemy.py:1776-1781)
    if not annihilate:
        assert 'lambda_sterics=1.0;' in aa['expression']
    for f in sterics:
        assert f['lrc'] == (not disable_lrc) and f['per_particle'] == ['sigma', 'epsilon'] and f['n_particles'] == N
        assert (f['globals']['softcore_alpha'], f['globals']['softcore_a'], f['globals']['softcore_b'], f['globals']['softcore_c']) == (alpha, a, b, cc)
    nb = desc[''][0]
    assert nb['type'] == 'NonbondedForce' and all(e == 0.0 for e in nb['eps'][:N_ALCH]) and all(e > 0.0 for e in nb['eps'][N_ALCH:])
    assert nb['dispersion'] is True


def test_package_factory_carries_the_same_switches():
    """The product-side factory exposes what the reference factory consumed above."""
    from openmmtools_b200 import alchemy, testsystems
    fluid = testsystems.LennardJonesFluid(nparticles=64)
    for annihilate, disable_lrc, (alpha, a, b, cc) in CONFIGS:
        region = alchemy.AlchemicalRegion(alchemical_atoms=range(6), annihilate_sterics=annihilate, softcore_alpha=alpha,
                                          softcore_a=a, softcore_b=b, softcore_c=cc)
        asys = alchemy.AbsoluteAlchemicalFactory(disable_alchemical_dispersion_correction=disable_lrc).create_alchemical_system(
            fluid.system, region)
        assert list(np.nonzero(asys.alchemical_mask())[0]) == list(range(6))
        assert bool(asys.annihilate_sterics) == annihilate
        assert (asys.softcore_alpha, asys.softcore_a, asys.softcore_b, asys.softcore_c) == (alpha, a, b, cc)
        assert bool(asys.alchemical_dispersion_correction) == (not disable_lrc)
    # the reference's defaults (alchemy.py:417-429, :626-628)
    r = alchemy.AlchemicalRegion(alchemical_atoms=[0])
    assert (r.annihilate_sterics, r.annihilate_electrostatics, r.softcore_alpha, r.softcore_a, r.softcore_b, r.softcore_c) == \
           (False, True, 0.5, 1, 1, 6)
