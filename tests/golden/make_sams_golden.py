#!/usr/bin/env python
"""Golden vectors for the SAMS state update from the REAL reference code (build container only).

``SAMSSampler._global_jump``, ``_update_logZ_estimates``, ``_update_stage``, ``_update_log_weights`` and
``MultiStateSampler._neighborhood`` are lifted by AST from /root/reference/openmmtools/multistate/{sams,multistatesampler}.py
(the modules cannot be imported: no OpenMM) and driven on synthetic energies with numpy's global RandomState seeded.
Output: tests/golden/sams_golden.npz
"""
import ast, os, sys, types
import numpy as np
from scipy.special import logsumexp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from energy_models import pseudo_normal


def lift(path, cls_name, names):
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name][0]
    out = {}
    for n in cls.body:
        if isinstance(n, ast.FunctionDef) and n.name in names:
            n.decorator_list = []
            ns = {'np': np, 'logsumexp': logsumexp,
                  'logger': types.SimpleNamespace(debug=lambda *a, **k: None)}
            exec(ast.unparse(n), ns)
            out[n.name] = ns[n.name]
    return out


sams = lift('/root/reference/openmmtools/multistate/sams.py', 'SAMSSampler',
            ['_global_jump', '_update_logZ_estimates', '_update_stage', '_update_log_weights'])
base = lift('/root/reference/openmmtools/multistate/multistatesampler.py', 'MultiStateSampler', ['_neighborhood'])


class Fake:
    pass


for k, v in {**sams, **base}.items():
    setattr(Fake, k, v)
Fake._state_histogram = property(lambda self: self._cached_state_histogram)


def energies_for(it, K, M, key):
    """u[k, l] of iteration `it`: harmonic-ish ladder plus iteration dependent noise (exact arithmetic only)."""
    x = pseudo_normal(K, key * 1000 + it)
    mu = 0.7 * np.arange(M, dtype=np.float64)
    return 0.5 * (3.0 * x[:, None] - mu[None, :]) ** 2 * 0.1 + 0.05 * mu[None, :]


out = {}
cases = []
for (K, M, seed, update_stages, method, criteria, gamma0) in [
        (1, 8, 0, 'two-stage', 'rao-blackwellized', 'logZ-flatness', 1.0),
        (3, 12, 1, 'two-stage', 'optimal', 'minimum-visits', 1.0),
        (2, 16, 7, 'one-stage', 'rao-blackwellized', 'logZ-flatness', 0.5),
        (4, 6, 3, 'two-stage', 'rao-blackwellized', 'histogram-flatness', 2.0)]:
    f = Fake()
    f.n_replicas, f.n_states, f.locality = K, M, None
    f.update_stages, f.weight_update_method, f.flatness_criteria = update_stages, method, criteria
    f.flatness_threshold, f.gamma0 = 0.2, gamma0
    f.log_target_probabilities = np.zeros(M) - np.log(M)
    f._logZ = np.zeros(M)
    f._t0 = 0
    f._stage = 1 if update_stages == 'one-stage' else 0
    f._cached_state_histogram = np.zeros(M, dtype=int)
    f._reporter = types.SimpleNamespace(write_online_analysis_data=lambda *a, **k: None)
    f._replica_thermodynamic_states = np.linspace(0, M - 1, K, dtype=int) if K > 1 else np.zeros(1, dtype=int)
    f._n_accepted_matrix = np.zeros((M, M), np.int64)
    f._n_proposed_matrix = np.zeros((M, M), np.int64)
    f._neighborhoods = np.ones((K, M), np.int8)
    f._update_log_weights()
    np.random.seed(seed)
    T = 60
    hist = dict(states=[], logZ=[], log_weights=[], stage=[], t0=[], nacc=[], nprop=[])
    for it in range(1, T + 1):
        f._iteration = it
        f._energy_thermodynamic_states = energies_for(it, K, M, seed + 17)
        f._n_accepted_matrix[:] = 0; f._n_proposed_matrix[:] = 0
        logP = np.zeros((K, M))
        f._global_jump(logP)
        f._update_logZ_estimates(logP)
        f._update_log_weights()
        # what _report_iteration_items does to the histogram (sams.py:385-393)
        st, cnt = np.unique(f._replica_thermodynamic_states, return_counts=True)
        f._cached_state_histogram[st] += cnt
        for k2, v in (('states', f._replica_thermodynamic_states), ('logZ', f._logZ), ('log_weights', f.log_weights),
                      ('stage', f._stage), ('t0', f._t0), ('nacc', f._n_accepted_matrix), ('nprop', f._n_proposed_matrix)):
            hist[k2].append(np.array(v).copy())
    tag = 'sams_K%d_M%d_s%d_%s_%s_%s' % (K, M, seed, update_stages, method, criteria)
    for k2, v in hist.items():
        out[tag + '_' + k2] = np.array(v)
    out[tag + '_gamma0'] = gamma0
    cases.append(tag)
out['cases'] = np.array(cases)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sams_golden.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst), cases)
