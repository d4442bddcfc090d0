"""The CPU oracle's mixing/RNG restatement against golden vectors captured from the REAL reference
(numba-jitted ReplicaExchangeSampler._mix_all_replicas_numba and the numpy-RandomState neighbour scheme;
tests/golden/make_mixing_golden.py)."""
import os
import numpy as np
import pytest
from oracle import oracle
from energy_models import energies

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'mixing_golden.npz'))


def digest(mat):
    m = mat.astype(np.uint64).ravel()
    w = (np.arange(m.size, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(1))
    return np.array([m.sum(), (m * m).sum(), (m * w).sum()], dtype=np.uint64)


def check_counts(key, mat):
    if mat.shape[0] >= 100:
        assert np.array_equal(G[key + '_digest'], digest(mat)), key
    else:
        assert np.array_equal(G[key], mat), key


@pytest.mark.parametrize('seed', [0, 1, 1234, 2**32 - 1])
def test_rng_known_answers(seed):
    for n in (1, 2, 3, 5, 64, 100, 256, 1000):
        mt = oracle.MT(seed)
        got = [mt.randint(n) for _ in range(16)]
        assert got == list(G[f'randint_s{seed}_n{n}']), (seed, n)
    mt = oracle.MT(seed)
    got = [mt.rand() for _ in range(16)]
    assert got == list(G[f'rand_s{seed}'])


def test_randint_one_does_not_consume_a_word():
    mt = oracle.MT(7)
    [mt.randint(1) for _ in range(3)]
    assert [mt.randint(256) for _ in range(4)] == list(G['interleave_n1_then_256_s7'])


def parse(tag):
    _, K, s, model = tag.split('_')
    return int(K[1:]), int(s[1:]), model


@pytest.mark.parametrize('tag', [str(t) for t in G['all_cases']])
def test_swap_all_matches_numba(tag):
    K, seed, model = parse(tag)
    u = energies(model, K, K * 1000 + (seed % 1000))
    st = np.arange(K, dtype=np.int64)
    mt = oracle.MT(seed)
    for call, nswap in ((1, K**3), (2, K**3), (3, 777)):
        nacc = np.zeros((K, K), np.int64); nprop = np.zeros((K, K), np.int64)
        oracle.mix_swap_all(mt, nswap, st, u, nacc, nprop)
        assert np.array_equal(st, G[f'{tag}_perm{call}']), (tag, call)
        check_counts(f'{tag}_nacc{call}', nacc)
        check_counts(f'{tag}_nprop{call}', nprop)


@pytest.mark.parametrize('tag', [str(t) for t in G['nbr_cases']])
def test_swap_neighbors_matches_numpy(tag):
    K, seed, model = parse(tag)
    u = energies(model, K, K * 77 + seed % 1000)
    st = np.arange(K, dtype=np.int64)
    mt = oracle.MT(seed)
    for it in range(6):
        nacc = np.zeros((K, K), np.int64); nprop = np.zeros((K, K), np.int64)
        oracle.mix_swap_neighbors(mt, st, u, nacc, nprop)
        assert np.array_equal(st, G[f'{tag}_perms'][it]), (tag, it)
    check_counts(f'{tag}_nacc_last', nacc)
    check_counts(f'{tag}_nprop_last', nprop)
