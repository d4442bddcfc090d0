#!/usr/bin/env python
"""Golden data for small host-side rules of the samplers, from the REAL reference code (build container only):

* ``MultiStateSampler._default_initial_thermodynamic_states`` (multistatesampler.py:1116-1143), lifted by AST;
* the temperature ladder of ``ParallelTemperingSampler.create`` (paralleltempering.py:109-173): the method is lifted
  and run with a recording base class, so the ladder is whatever the reference code computes.

Output: tests/golden/hostlogic_golden.npz
"""
import ast, copy, logging, os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from openmmtools_b200 import unit as u


def lift(path, cls_name, name):
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    fn.decorator_list = []
    return fn


if __name__ == '__main__':
    out = {}
    # ---- initial state assignment
    fn = lift('/root/reference/openmmtools/multistate/multistatesampler.py', 'MultiStateSampler', '_default_initial_thermodynamic_states')
    ns = {'np': np}
    exec(ast.unparse(fn), ns)
    f = ns['_default_initial_thermodynamic_states']
    cases = [(1, 1), (3, 3), (5, 1), (5, 2), (8, 3), (16, 5), (3, 7), (4, 9), (6, 6), (64, 10), (2, 5)]
    for n_thermo, n_sampler in cases:
        out['init_%d_%d' % (n_thermo, n_sampler)] = np.asarray(f(None, [None] * n_thermo, [None] * n_sampler))
    out['init_cases'] = np.array(cases)

    # ---- parallel tempering ladder
    sys.modules['openmm'] = types.ModuleType('openmm'); sys.modules['openmm'].unit = u; sys.modules['openmm.unit'] = u
    fn = lift('/root/reference/openmmtools/multistate/paralleltempering.py', 'ParallelTemperingSampler', 'create')

    class FakeState:
        def __init__(self): self.temperature = None

    class Base:
        def create(self, thermodynamic_states, sampler_states, storage=None, **kwargs):
            self.created = [s.temperature for s in thermodynamic_states]

    ns = {'np': np, 'copy': copy, 'logger': logging.getLogger('pt'), 'states': types.SimpleNamespace(ThermodynamicState=FakeState),
          'Base': Base}
    src = 'class PT(Base):\n' + '\n'.join('    ' + l for l in ast.unparse(fn).splitlines())
    exec(src, ns)
    ladders = [(300.0, 600.0, 128), (273.15, 373.15, 6), (300.0, 310.0, 2)]
    for tmin, tmax, n in ladders:
        pt = ns['PT']()
        pt.create(FakeState(), [None], storage=None, min_temperature=tmin * u.kelvin, max_temperature=tmax * u.kelvin, n_temperatures=n)
        out['pt_%g_%g_%d' % (tmin, tmax, n)] = np.array([float(u.to_md(t)) for t in pt.created])
    out['pt_cases'] = np.array(ladders)
    dst = os.path.join(HERE, 'hostlogic_golden.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, os.path.getsize(dst))
    print(out['init_8_3'], out['init_3_7'], out['pt_273.15_373.15_6'])
