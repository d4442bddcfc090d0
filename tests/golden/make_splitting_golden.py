#!/usr/bin/env python
"""What the reference's LangevinIntegrator says about a list of splitting strings (build container only):
``integrators.py`` runs unmodified on the recording CustomIntegrator of make_integrator_golden.py; for every string we
store whether construction succeeds, the exception type if not, and the R/V/O counts the integrator derived
(``_ORV_counts``).  Output: tests/golden/splitting_golden.json"""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_integrator_golden import load_reference_integrators
from openmmtools_b200 import unit as u

STRINGS = ['V R O R V', 'O V R V O', 'R V O', 'V R R O R R V', 'O R V R O', 'V R V', 'R O', 'V O', 'V R O', ' V  R O R V ', 'VRORV',
           'V R X R V', 'V0 R O R V0', 'V0 V1 R O R V1 V0', 'V32 R O R V32', 'Va R O', '{ V R O R V }', 'O { V R V } O', '{ V R O', 'V R O }', '',
           'R R R O V', 'O O V R', 'v r o']

if __name__ == '__main__':
    mod = load_reference_integrators()
    out = {}
    for s in STRINGS:
        try:
            integ = mod.LangevinIntegrator(temperature=300 * u.kelvin, collision_rate=1 / u.picoseconds, timestep=1 * u.femtoseconds, splitting=s)
            out[s] = {'ok': True, 'counts': {k: int(v) for k, v in integ._ORV_counts.items()}, 'mts': bool(integ._mts),
                      'metropolized': bool(integ._metropolized_integrator)}
        except Exception as e:   # noqa: the type is the datum
            out[s] = {'ok': False, 'error': type(e).__name__, 'message': str(e)[:120]}
        print(repr(s), out[s])
    json.dump(out, open(os.path.join(HERE, 'splitting_golden.json'), 'w'), indent=1, sort_keys=True)
