"""Lane-level CPU model of k_mix_walk2c's control logic (openmmtools_b200/csrc/rx_walk2c.cuh): candidate coordinates
(the words numba's rejection-sampling randint accepts, numbered in stream order), one record per candidate index
{i, j, back-mask, the attempt's uniform, k = candidates among the uniform's two raw words}, a window of 32 candidate
indices, the visited chain for hops {2, 3, 4} resolved with stride-2 add-carry masks in grid phases, forced stops,
staleness, prefix commit and the hand-over to the plain loop at a word position.  The model must reproduce the sequential
reference loop (replicaexchange.py:321-349 with numba's randint, numba/_random.c) attempt by attempt."""
import sys, math, os, random
import numpy as np
import pytest
sys.path.insert(0, os.path.dirname(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
from energy_models import energies
from test_walk_any_logic_model import reference, uniform, popc

M32 = 0xffffffff


def chain(Gw, K1w, K2w):
    """w2c_chain, operation for operation."""
    NG = ~Gw & M32
    T = NG & K1w; F = NG & K2w
    XgE = F & 0x55555555; XgO = F & 0xAAAAAAAA
    XpE = (XgE | (XgE << 1)) & M32; XpO = (XgO | (XgO << 1)) & M32
    stE = XgE & ~(XgE << 2) & M32; stO = XgO & ~(XgO << 2) & M32
    sEa = (XpE + (stE & 0x11111111)) & M32; sEb = (XpE + (stE & 0x44444444)) & M32
    sOa = (XpO + (stO & 0x22222222)) & M32; sOb = (XpO + (stO & 0x88888888)) & M32
    VE = 0x55555555 & ~(((sEa ^ XpE) & 0x44444444) | ((sEb ^ XpE) & 0x11111111)) & M32
    VO = 0xAAAAAAAA & ~(((sOa ^ XpO) & 0x88888888) | ((sOb ^ XpO) & 0x22222222)) & M32
    Vc, Xc, gc, Vn, Xn, gn = VE, XpE, 0x55555555, VO, XpO, 0xAAAAAAAA
    st = 1; V = 0
    while True:     # one grid phase per visited T: from position st on the current grid up to its first visited T
        R = ((Xc + st) ^ Xc) & gc & M32
        Vg = (Vc ^ (R if (st & ~Vc & M32) else 0)) & ~((st - 1) & M32) & M32
        Tv = Vg & T; t = Tv & -Tv & M32
        V |= Vg & (((t << 1) - 1) & M32)
        st = (t << 3) & M32
        if st == 0: break
        Vc, Xc, gc, Vn, Xn, gn = Vn, Xn, gn, Vc, Xc, gc
    Cf = V & (0x80000000 | ((T | F) & 0x40000000) | (F & 0x20000000))   # positions whose hop ends beyond position 32
    return V, Cf


def test_chain_masks_reproduce_the_sequential_chain():
    rng = random.Random(3)
    for _ in range(60000):
        pg = rng.choice([0.1, 0.3, 0.5, 0.69, 0.9, 1.0]); pk = rng.choice([0.5, 0.78, 1.0])
        G = K1 = K2 = 0
        for b in range(32):
            if rng.random() < pg: G |= 1 << b
            k = (rng.random() < pk) + (rng.random() < pk)
            if k == 1: K1 |= 1 << b
            if k == 2: K2 |= 1 << b
        V, Cf = chain(G, K1, K2)
        pos = 0; Vs = 0
        while pos < 32:
            Vs |= 1 << pos
            pos += 2 if (G >> pos) & 1 else 2 + ((K1 >> pos) & 1) + 2 * ((K2 >> pos) & 1)
        assert V == Vs
        last = Vs.bit_length() - 1           # the chain's last visited position either is the forced stop or hops to 32 exactly
        hop = 2 if (G >> last) & 1 else 2 + ((K1 >> last) & 1) + 2 * ((K2 >> last) & 1)
        assert Cf == ((1 << last) if last + hop > 32 else 0)


def run(K, nswap, model, seed):
    u = energies(model, K, 99).tolist()
    nbits = (K - 1).bit_length()
    mask = (1 << nbits) - 1
    words = [int(x) for x in np.random.RandomState(seed)._bit_generator.random_raw(12 * nswap + 6000).astype(np.uint64)]
    perm_ref, head_ref, log_ref = reference(K, nswap, u, words, mask)
    # ---- pre-pass (k_cand_count / scan / scatter / records)
    cpos = [p for p, x in enumerate(words) if (x & mask) < K]
    ncand = len(cpos)
    val = [words[p] & mask for p in cpos]
    nslots = ncand - 3

    def rec(c):
        i, j = val[c], val[c + 1]
        a = cpos[c + 1] + 1
        k = (cpos[c + 2] <= a + 1) + (cpos[c + 3] <= a + 1)
        bm = 0
        for b in range(31):
            cc = c - 1 - b
            if cc >= 0 and (val[cc] in (i, j) or val[cc + 1] in (i, j)): bm |= 1 << (31 - b)
        return i, j, k, a, bm
    # ---- walker
    perm = list(range(K)); h = 0; rem = nswap; log = []; rounds = 0
    h_end = nslots - 560 if nslots >= 1200 else 0
    while rem >= 130 and nslots >= 1200 and h + 99 <= h_end:
        rounds += 1
        lanes = []
        for w in range(32):
            i, j, k, a, bm = rec(h + w)
            si, sj = perm[i], perm[j]
            logp = -(u[i][sj] + u[j][si]) + u[i][si] + u[j][sj]
            ge0 = logp >= 0
            acc = ge0 or uniform(words, a) < math.exp(logp)
            lanes.append(dict(i=i, j=j, k=k, si=si, sj=sj, ge0=ge0, acc=acc, bm=bm))
        ballot = lambda f: sum((1 << l) for l in range(32) if f(lanes[l]))
        Gw = ballot(lambda x: x['ge0']); Aw = ballot(lambda x: x['acc'] and x['i'] != x['j'])
        K1w = ballot(lambda x: x['k'] == 1); K2w = ballot(lambda x: x['k'] == 2)
        V, Cf = chain(Gw, K1w, K2w)
        VA = V & Aw
        C = 0
        for w in range(32):
            earlier = (VA << (32 - w)) & M32 if w else 0
            if earlier & lanes[w]['bm']: C |= 1 << w
        Cw = (C | Cf) & V
        assert not (Cw & 1)
        low = Cw & -Cw
        below = (low - 1) & M32              # Cw == 0: the whole window commits, the next round starts at position 32
        cm = V & below
        for w in range(32):
            if (cm >> w) & 1:
                x = lanes[w]; log.append((x['si'], x['sj'], x['acc']))
        for w in range(32):
            x = lanes[w]
            if (cm >> w) & 1 and x['acc'] and x['i'] != x['j']:
                perm[x['i']] = x['sj']; perm[x['j']] = x['si']
        h += popc(below); rem -= popc(cm)
    # ---- the plain loop finishes from the word position of candidate h (k_mix_walk_serial)
    p = cpos[h] if h > 0 else 0
    while rem > 0:
        while True:
            i = words[p] & mask; p += 1
            if i < K: break
        while True:
            j = words[p] & mask; p += 1
            if j < K: break
        si, sj = perm[i], perm[j]
        logp = -(u[i][sj] + u[j][si]) + u[i][si] + u[j][sj]
        if logp >= 0: acc = True
        else:
            acc = uniform(words, p) < math.exp(logp); p += 2
        log.append((si, sj, acc))
        if acc: perm[i], perm[j] = sj, si
        rem -= 1
    assert log == log_ref
    assert perm == perm_ref and p == head_ref
    return rounds


@pytest.mark.parametrize('K,model,nswap,seed', [(3, 'flat', 3000, 1), (5, 'normal', 3000, 2), (6, 'zeros', 2000, 3),
                                                (12, 'flat', 4000, 4), (100, 'normal', 6000, 5), (65, 'flat', 5000, 6),
                                                (127, 'ladder', 4000, 7), (129, 'normal', 4000, 8), (200, 'ladder', 5000, 9),
                                                (255, 'flat', 4000, 10)])
def test_model_reproduces_reference(K, model, nswap, seed):
    rounds = run(K, nswap, model, seed)
    assert 0 < rounds < nswap        # the window commits more than one attempt per round (few at tiny K: conflicts)
