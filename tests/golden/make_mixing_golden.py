#!/usr/bin/env python
"""Generate golden vectors for the swap-all / swap-neighbors mixing kernels from the REAL reference code.

Run in the build container only (needs /root/reference); the GPU box never runs this.
The reference function ``ReplicaExchangeSampler._mix_all_replicas_numba`` is lifted by AST from
``/root/reference/openmmtools/multistate/replicaexchange.py`` (the module itself cannot be imported:
``import openmm`` fails), compiled with numba.njit, and driven with a seeded numba generator.
``_attempt_swap`` / ``_mix_neighboring_replicas`` are lifted the same way and driven with numpy's
global RandomState (which is what the reference uses for that scheme).

Output: tests/golden/mixing_golden.npz  (committed).
"""
import ast, os, sys, textwrap, math
import numpy as np
from numba import njit

REF = '/root/reference/openmmtools/multistate/replicaexchange.py'
src = open(REF).read()
tree = ast.parse(src)
cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'ReplicaExchangeSampler'][0]
funcs = {n.name: n for n in cls.body if isinstance(n, ast.FunctionDef)}

def lift(name):
    node = funcs[name]
    node.decorator_list = []
    code = ast.unparse(node)
    ns = {'np': np, 'math': math, 'logger': type('L', (), {'debug': staticmethod(lambda *a, **k: None)})}
    exec(code, ns)
    return ns[name]

mix_numba = njit(lift('_mix_all_replicas_numba'))
attempt_swap = lift('_attempt_swap')
mix_neighbors = lift('_mix_neighboring_replicas')

@njit
def numba_seed(s):
    np.random.seed(s)

@njit
def numba_draws(n, cnt):
    out = np.empty(cnt, np.int64)
    for i in range(cnt):
        out[i] = np.random.randint(n)
    return out

@njit
def numba_rand(cnt):
    out = np.empty(cnt, np.float64)
    for i in range(cnt):
        out[i] = np.random.rand()
    return out

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from energy_models import energies

def digest(mat):
    """Order-sensitive digest of a count matrix (used instead of the full matrix for K >= 100)."""
    m = mat.astype(np.uint64).ravel()
    w = (np.arange(m.size, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))
    return np.array([m.sum(), (m * m).sum(), (m * w).sum()], dtype=np.uint64)

def put_counts(out, key, mat):
    if mat.shape[0] >= 100:
        out[key + '_digest'] = digest(mat)
    else:
        out[key] = mat.astype(np.int32)

class FakeSampler:
    """Just enough of ReplicaExchangeSampler for the lifted _mix_neighboring_replicas/_attempt_swap."""
    def __init__(self, K, u):
        self.n_replicas = K
        self._replica_thermodynamic_states = np.arange(K, dtype=np.int64)
        self._energy_thermodynamic_states = u
        self._n_accepted_matrix = np.zeros((K, K), np.int64)
        self._n_proposed_matrix = np.zeros((K, K), np.int64)
    _attempt_swap = attempt_swap
    _mix_neighboring_replicas = mix_neighbors

out = {}
# --- RNG known answers
for seed in (0, 1, 1234, 2**32 - 1):
    for n in (1, 2, 3, 5, 64, 100, 256, 1000):
        numba_seed(seed)
        out[f'randint_s{seed}_n{n}'] = numba_draws(n, 16)
    numba_seed(seed)
    out[f'rand_s{seed}'] = numba_rand(16)
# interleaved: does randint(1) consume a word?
numba_seed(7); a = numba_draws(1, 3); b = numba_draws(256, 4)
out['interleave_n1_then_256_s7'] = b
numba_seed(7); out['plain_256_s7'] = numba_draws(256, 8)

# --- swap-all chains
cases = []
for K in (2, 3, 5, 16, 33, 64, 100, 128):
    for seed in (0, 1, 1234, 2**32 - 1):
        for model in ('zeros', 'normal', 'ladder', 'flat'):
            if K >= 100:
                continue
            cases.append((K, seed, model))
cases += [(100, 0, 'ladder'), (100, 1234, 'flat'), (128, 0, 'normal'), (128, 1234, 'zeros')]
cases.append((256, 1234, 'ladder'))
cases.append((256, 0, 'flat'))
for (K, seed, model) in cases:
    u = energies(model, K, K * 1000 + (seed % 1000))
    st = np.arange(K, dtype=np.int64)
    nacc = np.zeros((K, K), np.int64); nprop = np.zeros((K, K), np.int64)
    numba_seed(seed)
    mix_numba(K ** 3, K, st, u, nacc, nprop)
    tag = f'all_K{K}_s{seed}_{model}'
    out[tag + '_perm1'] = st.copy()
    put_counts(out, tag + '_nacc1', nacc); put_counts(out, tag + '_nprop1', nprop)
    # second consecutive call: stream continues, perm carried over, counts re-zeroed (as _mix_replicas does)
    nacc[:] = 0; nprop[:] = 0
    mix_numba(K ** 3, K, st, u, nacc, nprop)
    out[tag + '_perm2'] = st.copy()
    put_counts(out, tag + '_nacc2', nacc); put_counts(out, tag + '_nprop2', nprop)
    # a reduced-attempt call (nswap != K^3), stream continues
    nacc[:] = 0; nprop[:] = 0
    mix_numba(777, K, st, u, nacc, nprop)
    out[tag + '_perm3'] = st.copy()
    put_counts(out, tag + '_nacc3', nacc); put_counts(out, tag + '_nprop3', nprop)
out['all_cases'] = np.array([f'all_K{K}_s{seed}_{model}' for (K, seed, model) in cases])

# --- swap-neighbors chains (numpy global RandomState)
ncases = []
for K in (2, 3, 5, 16, 64, 256):
    for seed in (0, 1234):
        for model in ('zeros', 'ladder', 'flat'):
            u = energies(model, K, K * 77 + seed % 1000)
            fs = FakeSampler(K, u)
            np.random.seed(seed)
            perms = []
            for it in range(6):
                fs._n_accepted_matrix[:] = 0; fs._n_proposed_matrix[:] = 0
                fs._mix_neighboring_replicas()
                perms.append(fs._replica_thermodynamic_states.copy())
            tag = f'nbr_K{K}_s{seed}_{model}'
            out[tag + '_perms'] = np.array(perms)
            put_counts(out, tag + '_nacc_last', fs._n_accepted_matrix)
            put_counts(out, tag + '_nprop_last', fs._n_proposed_matrix)
            ncases.append(tag)
out['nbr_cases'] = np.array(ncases)

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mixing_golden.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst), 'bytes;', len(cases), 'swap-all cases,', len(ncases), 'swap-neighbors cases')
