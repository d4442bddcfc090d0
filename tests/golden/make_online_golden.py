#!/usr/bin/env python
"""Golden vectors for the per-iteration online free-energy estimate from the REAL reference code (build container only).

``MultiStateSampler._online_analysis`` and ``_neighborhood`` are lifted by AST from
/root/reference/openmmtools/multistate/multistatesampler.py (the module cannot be imported: no OpenMM) and driven on
synthetic energies and replica states.  Output: tests/golden/online_golden.npz
"""
import ast, os, sys, types
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from energy_models import ladder_energies as energies_for, ladder_states as states_for


class _Timer:
    def start(self, *a): pass
    def stop(self, *a): pass


def lift(path, cls_name, names):
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name][0]
    out = {}
    for n in cls.body:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            n.decorator_list = []
            ns = {'np': np, 'utils': types.SimpleNamespace(Timer=_Timer),
                  'logger': types.SimpleNamespace(debug=lambda *a, **k: None)}
            exec(ast.unparse(n), ns)
            out[n.name] = ns[n.name]
    return out


class Fake:
    pass


for k, v in lift('/root/reference/openmmtools/multistate/multistatesampler.py', 'MultiStateSampler',
                 ['_online_analysis', '_neighborhood']).items():
    setattr(Fake, k, v)


out, cases = {}, []
for (K, M, locality, key) in [(4, 4, None, 3), (6, 9, None, 5), (5, 12, 2, 8), (3, 7, 1, 11)]:
    f = Fake()
    f.n_states, f.locality = M, locality
    f._last_mbar_f_k = None
    f._reporter = types.SimpleNamespace(write_online_data_dynamic_and_static=lambda *a, **k: None)
    hist, errs = [], []
    for it in range(1, 41):
        f._iteration = it
        f._energy_thermodynamic_states = energies_for(it, K, M, key)
        f._replica_thermodynamic_states = states_for(it, K, M, key)
        errs.append(f._online_analysis())
        hist.append(f._last_mbar_f_k.copy())
    tag = 'online_K%d_M%d_loc%s' % (K, M, locality)
    out[tag + '_f_k'] = np.array(hist)
    out[tag + '_key'] = key
    assert all(np.isinf(e) for e in errs)
    cases.append(tag)
out['cases'] = np.array(cases)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'online_golden.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst), cases)
