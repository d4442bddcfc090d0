#!/usr/bin/env python
"""Where a real OpenMM + openmmtools install exists: dump Reference-platform energies (and forces) of the EXACT inputs
this repository's parity tests use, so the "parity UNPINNED" status of the energy function can be
closed by diffing against tests/golden/openmm_lj_reference.npz.

UNTESTED in the build container (OpenMM is not installable there: no wheel, no network).  It uses only public API:
openmmtools.testsystems.LennardJonesFluid, alchemy.AbsoluteAlchemicalFactory / AlchemicalState,
states.ThermodynamicState / CompoundThermodynamicState / SamplerState, cache.ContextCache(platform=Reference)
(the idiom of /root/reference/openmmtools/tests/test_sampling.py:111-113).

  python tools/export_openmm_oracle.py            # writes tests/golden/openmm_lj_reference.npz
  python -m pytest tests/test_openmm_golden.py    # (skipped when the file is absent)
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import openmm
    from openmm import unit
    from openmmtools import testsystems, alchemy, states, cache
    from helpers import lj_setup                      # the same jittered configurations the GPU tests use

    platform = openmm.Platform.getPlatformByName('Reference')
    ctx_cache = cache.ContextCache(platform=platform)
    out = {}
    for n_alch, disable_lrc, annihilate in ((10, True, False), (10, False, False), (10, True, True)):
        fluid = testsystems.LennardJonesFluid(nparticles=512)
        factory = alchemy.AbsoluteAlchemicalFactory(disable_alchemical_dispersion_correction=disable_lrc)
        region = alchemy.AlchemicalRegion(alchemical_atoms=range(n_alch), annihilate_sterics=annihilate)
        asys = factory.create_alchemical_system(fluid.system, region)
        astate = alchemy.AlchemicalState.from_system(asys)
        lambdas = np.linspace(1.0, 0.0, 16)
        s = lj_setup(N=512, n_alch=n_alch, seed=3)
        x = s['x']
        tag = 'lj512_a%d_lrc%d_ann%d' % (n_alch, int(not disable_lrc), int(annihilate))
        u = np.zeros(len(lambdas))
        for l, lam in enumerate(lambdas):
            cstate = states.CompoundThermodynamicState(states.ThermodynamicState(asys, 300 * unit.kelvin), [astate])
            cstate.lambda_sterics = float(lam)
            cstate.lambda_electrostatics = float(lam)
            sstate = states.SamplerState(x * unit.nanometer, box_vectors=asys.getDefaultPeriodicBoxVectors())
            context, _ = ctx_cache.get_context(cstate)
            sstate.apply_to_context(context)
            u[l] = cstate.reduced_potential(context)
            if l in (0, len(lambdas) // 2):
                st = context.getState(getForces=True, getEnergy=True)
                out['%s_forces_l%d' % (tag, l)] = st.getForces(asNumpy=True).value_in_unit(unit.kilojoule_per_mole / unit.nanometer)
                out['%s_U_l%d' % (tag, l)] = st.getPotentialEnergy().value_in_unit(unit.kilojoule_per_mole)
        out[tag + '_u'] = u
        out[tag + '_lambdas'] = lambdas
        out[tag + '_x'] = x
    out['openmm_version'] = np.array(openmm.version.version)
    dst = os.path.join(ROOT, 'tests', 'golden', 'openmm_lj_reference.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst)


if __name__ == '__main__':
    main()
