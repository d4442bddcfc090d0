"""The drop-in's constructors and methods accept the reference's arguments, in the reference's order, with the
reference's defaults (tests/golden/signatures_golden.json, read from the reference's AST by
tests/golden/make_signatures_golden.py).  Extra keyword arguments of this package (seed, communicator, ...) come after."""
import inspect
import json
import logging
import os
import numpy as np
import pytest
from openmmtools_b200 import mcmc, multistate, states, alchemy, testsystems, cache, unit
from openmmtools_b200.multistate import multistatereporter

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'signatures_golden.json')))
WHERE = {
    'mcmc.py': mcmc, 'multistate/multistatesampler.py': multistate, 'multistate/replicaexchange.py': multistate,
    'multistate/paralleltempering.py': multistate, 'multistate/sams.py': multistate,
    'multistate/multistatereporter.py': multistatereporter, 'states.py': states, 'alchemy/alchemy.py': alchemy,
    'testsystems.py': testsystems, 'cache.py': cache,
}
# arguments of reference features this build does not carry (each raises NotImplementedError/ValueError when used)
NOT_CARRIED = {
    'testsystems.py:LennardJonesFluid.__init__': {'shift', 'lattice', 'charge', 'ewaldErrorTolerance'},
}


def _value(src):
    return eval(src, {'unit': unit, 'np': np, 'logging': logging, 'frozenset': frozenset,
                      'DEFAULT_EWALD_ERROR_TOLERANCE': None})


def _same(a, b):
    if hasattr(a, 'unit') or hasattr(b, 'unit'):
        return np.allclose(np.asarray(unit.to_md(a), float), np.asarray(unit.to_md(b), float), rtol=1e-12)
    if isinstance(a, float) or isinstance(b, float):
        return a == b or (a != a and b != b)
    return a == b


@pytest.mark.parametrize('key', sorted(k for k in G if '.' in k.split(':')[1]))
def test_signature_matches_reference(key):
    path, qual = key.split(':')
    cls_name, fn_name = qual.split('.')
    cls = getattr(WHERE[path], cls_name)
    fn = getattr(cls, fn_name)
    params = [p for p in inspect.signature(fn).parameters.values() if p.name not in ('self', 'cls')]
    names = [p.name for p in params]
    skip = NOT_CARRIED.get(key, set())
    ref = [(n, d) for n, d in G[key]['args'] if n not in skip]
    has_kwargs = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params)
    pos = 0
    for n, d in ref:
        if n not in names:
            assert has_kwargs, '%s lacks argument %r' % (key, n)   # forwarded through **kwargs to a base class that has it
            continue
        i = names.index(n)
        assert i >= pos, '%s: argument %r out of the reference order' % (key, n)
        pos = i
        if d is not None:
            mine = params[i].default
            assert mine is not inspect.Parameter.empty, '%s: %r must have a default' % (key, n)
            assert _same(mine, _value(d)), '%s: default of %r is %r, reference %s' % (key, n, mine, d)


def test_alchemical_region_fields_and_defaults():
    ref = G['alchemy/alchemy.py:AlchemicalRegion']['args']
    r = alchemy.AlchemicalRegion()
    for name, default in ref:
        assert hasattr(r, name), name
        assert _same(getattr(r, name), eval(default)), (name, getattr(r, name), default)
