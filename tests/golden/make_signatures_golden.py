#!/usr/bin/env python
"""Constructor / method signatures of the reference classes the drop-in mirrors (build container only): argument
names, order and default-value source text, read from the reference's AST.  Output: tests/golden/signatures_golden.json"""
import ast, json, os

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/openmmtools/'
TARGETS = [
    ('mcmc.py', 'LangevinSplittingDynamicsMove', '__init__'), ('mcmc.py', 'LangevinDynamicsMove', '__init__'),
    ('multistate/multistatesampler.py', 'MultiStateSampler', '__init__'), ('multistate/multistatesampler.py', 'MultiStateSampler', 'create'),
    ('multistate/multistatesampler.py', 'MultiStateSampler', 'minimize'), ('multistate/multistatesampler.py', 'MultiStateSampler', 'equilibrate'),
    ('multistate/multistatesampler.py', 'MultiStateSampler', 'run'), ('multistate/multistatesampler.py', 'MultiStateSampler', 'extend'),
    ('multistate/replicaexchange.py', 'ReplicaExchangeSampler', '__init__'),
    ('multistate/paralleltempering.py', 'ParallelTemperingSampler', 'create'),
    ('multistate/sams.py', 'SAMSSampler', '__init__'),
    ('multistate/multistatereporter.py', 'MultiStateReporter', '__init__'),
    ('states.py', 'ThermodynamicState', '__init__'), ('states.py', 'SamplerState', '__init__'),
    ('states.py', 'CompoundThermodynamicState', '__init__'),
    ('alchemy/alchemy.py', 'AbsoluteAlchemicalFactory', '__init__'), ('alchemy/alchemy.py', 'AbsoluteAlchemicalFactory', 'create_alchemical_system'),
    ('alchemy/alchemy.py', 'AlchemicalState', 'from_system'),
    ('testsystems.py', 'LennardJonesFluid', '__init__'), ('testsystems.py', 'HarmonicOscillator', '__init__'),
    ('cache.py', 'ContextCache', '__init__'),
]
out = {}
for path, cls_name, fn_name in TARGETS:
    tree = ast.parse(open(REF + path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == fn_name][0]
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    kwonly = [(x.arg, None if d is None else ast.unparse(d)) for x, d in zip(a.kwonlyargs, a.kw_defaults)]
    out['%s:%s.%s' % (path, cls_name, fn_name)] = {
        'args': [[n, d] for n, d in zip(names, defaults) if n not in ('self', 'cls')], 'kwonly': kwonly,
        'vararg': a.vararg.arg if a.vararg else None, 'kwarg': a.kwarg.arg if a.kwarg else None, 'line': fn.lineno}
# AlchemicalRegion fields
tree = ast.parse(open(REF + 'alchemy/alchemy.py').read())
for n in tree.body:
    if isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == '_ALCHEMICAL_REGION_ARGS' for t in n.targets):
        ns = {}
        exec('import collections\n' + ast.unparse(n), ns)
        out['alchemy/alchemy.py:AlchemicalRegion'] = {'args': [[k, repr(v)] for k, v in ns['_ALCHEMICAL_REGION_ARGS'].items()]}
dst = os.path.join(HERE, 'signatures_golden.json')
json.dump(out, open(dst, 'w'), indent=1, sort_keys=True)
print('wrote', dst, len(out))
for k, v in out.items():
    print(k, [a[0] for a in v['args']])
