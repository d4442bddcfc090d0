"""If tests/golden/openmm_lj_reference.npz exists (made by tools/export_openmm_oracle.py on a machine with OpenMM), pin
the CPU oracle's energy function to OpenMM's Reference platform: 1e-5 relative on reduced potentials.  Skipped otherwise
(the energy parity is then 'unpinned')."""
import os
import numpy as np
import pytest

PATH = os.path.join(os.path.dirname(__file__), 'golden', 'openmm_lj_reference.npz')


@pytest.mark.skipif(not os.path.exists(PATH), reason='no OpenMM golden vectors (OpenMM is not installable in the build container)')
def test_oracle_energies_match_openmm_reference_platform():
    from oracle import oracle
    from openmmtools_b200 import testsystems, alchemy, _backend
    G = np.load(PATH)
    KB = 8.31446261815324e-3
    for n_alch, lrc, ann in ((10, 0, 0), (10, 1, 0), (10, 0, 1)):
        tag = 'lj512_a%d_lrc%d_ann%d' % (n_alch, lrc, ann)
        fluid = testsystems.LennardJonesFluid(nparticles=512)
        asys = alchemy.AbsoluteAlchemicalFactory(disable_alchemical_dispersion_correction=not lrc).create_alchemical_system(
            fluid.system, alchemy.AlchemicalRegion(alchemical_atoms=range(n_alch), annihilate_sterics=bool(ann)))
        L = asys.box_vectors[0, 0]
        osys = oracle.LJSystem(asys.sigma, asys.epsilon, asys.masses, asys.alchemical_mask(), (L, L, L), asys.cutoff,
                               asys.switching_distance, use_switch=True, annihilate_sterics=bool(ann))
        lambdas = G[tag + '_lambdas']
        off = np.array([_backend.lj_dispersion_correction(asys) + _backend.alchemical_dispersion_correction(asys, l) for l in lambdas])
        u = osys.energy_row(G[tag + '_x'], lambdas, np.full(len(lambdas), 1.0 / (KB * 300.0)), off)
        assert np.allclose(u, G[tag + '_u'], rtol=1e-5, atol=1e-6), tag
