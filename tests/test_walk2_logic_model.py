"""Lane-level CPU model of k_mix_walk2's control logic (openmmtools_b200/csrc/rx_walk2.cuh): fixed lane <-> slot mapping,
window rotated over the lanes, visited-chain bit arithmetic, conflict ballot with undecided lanes folded in, deferred
(rotated) commit with promotion of the lanes that left the window, no 33-slot advance, and the exact re-evaluation of a
round whose first lane the filter could not decide (on the device the undecided lanes take the exact decision and inject it
into the next fast round, run alone: the same decisions, resolved by the same arithmetic).  The model must reproduce the sequential reference loop (replicaexchange.py:321-349) attempt by attempt; the
filter's 'undecided' answers are injected at random (a decision is then arbitrary, as on the device)."""
import sys, math, random, os
import numpy as np
import pytest
sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
from energy_models import energies
M32 = 0xffffffff
def rotr(x, r): r &= 31; return ((x >> r) | (x << (32 - r))) & M32 if r else x
def popc(x): return bin(x).count('1')
def run(K, nswap, model, seed, p_und=0.0):
    u = energies(model, K, 99).tolist()
    words = np.random.RandomState(seed)._bit_generator.random_raw(4 * nswap + 4000).astype(np.uint64)
    words = [int(x) for x in words]
    mask = K - 1
    nslots = len(words) // 2
    # ---- reference
    perm = list(range(K)); p = 0; log_ref = []
    for t in range(nswap):
        i = words[p] & mask; j = words[p+1] & mask; si, sj = perm[i], perm[j]
        logp = -(u[i][sj] + u[j][si]) + u[i][si] + u[j][sj]; s = p // 2; p += 2
        if logp >= 0: acc = True
        else:
            r_ = ((words[p] >> 5) * 67108864.0 + (words[p+1] >> 6)) / 9007199254740992.0
            acc = r_ < math.exp(logp); p += 2
        log_ref.append((s, si, sj, acc))
        if acc: perm[i], perm[j] = sj, si
    perm_ref = perm; head_ref = p // 2
    # ---- records
    ij = [(words[2*s] & mask, words[2*s+1] & mask) for s in range(nslots)]
    def backmask(s):
        bm = 0; i, j = ij[s]
        for b in range(31):
            q = s - 1 - b
            if q >= 0 and (ij[q][0] in (i, j) or ij[q][1] in (i, j)): bm |= 1 << (31 - b)
        return bm
    bms = {}
    def rec(s):
        if s not in bms: bms[s] = backmask(s)
        return (ij[s][0], ij[s][1], bms[s])
    def evaluate(s, i, j, si, sj):
        logp = -(u[i][sj] + u[j][si]) + u[i][si] + u[j][sj]
        if logp >= 0: return True, True
        r_ = ((words[2*(s+1)] >> 5) * 67108864.0 + (words[2*(s+1)+1] >> 6)) / 9007199254740992.0
        return False, r_ < math.exp(logp)
    # ---- emulation
    perm = list(range(K)); h = 0; rem = nswap; log = []
    h_end = nslots - 161
    rng = random.Random(1)
    lanes = range(32)
    r = h & 31; w = [(l - h) & 31 for l in lanes]; sA = [h + w[l] for l in lanes]
    A = [rec(sA[l]) for l in lanes]; B = [rec(sA[l] + 32) for l in lanes]
    st = [(perm[A[l][0]], perm[A[l][1]]) for l in lanes]
    P = dict(mine=[False]*32, swaps=[False]*32, prom=[False]*32, entry=[None]*32, adv=0, n=0)
    def commit():
        nonlocal h, r, rem
        for l in lanes:
            if P['swaps'][l]:
                i, j, _ = A[l]; si, sj = st[l]; perm[i] = sj; perm[j] = si
        for l in lanes:
            if P['mine'][l]: log.append(P['entry'][l])
        h += P['adv']; r = (r + P['adv']) & 31; rem -= P['n']
        for l in lanes:
            if P['prom'][l]: A[l] = B[l]; sA[l] += 32
            st[l] = (perm[A[l][0]], perm[A[l][1]])
            w[l] = (w[l] - P['adv']) & 31
            B[l] = rec(sA[l] + 32)
    def resolve(ge, ac, und):
        G = sum(1 << l for l in lanes if ge[l]); Am = sum(1 << l for l in lanes if ac[l] and A[l][0] != A[l][1])
        Gw, Aw = rotr(G, r), rotr(Am, r)
        X = ~Gw & M32; starts = X & ~(X << 1) & M32; SE = starts & 0x55555555; SO = starts & 0xAAAAAAAA
        sumE = (X + SE); sumO = (X + SO); carryO = sumO > M32; sumE &= M32; sumO &= M32
        skip = (((sumE ^ X) & ~SE) & 0xAAAAAAAA) | (((sumO ^ X) & ~SO) & 0x55555555)
        V = ~skip & M32; VA = V & Aw
        Cb = 0
        for l in lanes:
            sh = 32 - w[l]; earlier = (VA << sh) & M32 if sh < 32 else 0
            if ((earlier & A[l][2]) | und[l]) != 0: Cb |= 1 << l
        # (the last window position does not commit an attempt whose uniform lies beyond the window)
        Cw = (rotr(Cb, r) | (X & 0x80000000)) & V; low = Cw & (-Cw & M32); below = (low - 1) & M32; cm = V & below
        adv = popc(below)
        assert not (Cw == 0 and carryO)
        for l in lanes:
            P['mine'][l] = bool((cm >> w[l]) & 1)
            P['swaps'][l] = P['mine'][l] and ac[l] and A[l][0] != A[l][1]
            P['prom'][l] = bool((below >> w[l]) & 1)
            P['entry'][l] = (sA[l], st[l][0], st[l][1], ac[l])
        P['adv'] = adv; P['n'] = popc(cm)
        return cm
    rounds = 0
    while True:
        P.update(mine=[False]*32, swaps=[False]*32, prom=[False]*32, adv=0, n=0)
        cm = 0; adv33 = 0
        while True:
            commit(); rounds += 1
            go = rem >= 97 and h + 66 <= h_end
            ge = [False]*32; ac = [False]*32; und = [0]*32
            for l in lanes:
                i, j, bm = A[l]; si, sj = st[l]
                assert sA[l] == h + w[l], (sA[l], h, w[l])
                assert (i, j) == ij[sA[l]]
                ge[l], ac[l] = evaluate(sA[l], i, j, si, sj)
                if rng.random() < p_und: und[l] = 1; ge[l] = rng.random() < .5; ac[l] = rng.random() < .5
            cm = resolve(ge, ac, und)
            adv33 = 0
            if not (go and cm != 0 and adv33 == 0): break
        commit()
        if adv33:
            for l in lanes:
                if w[l] == 31:
                    sA[l] += 32; A[l] = rec(sA[l]); B[l] = rec(sA[l] + 32); st[l] = (perm[A[l][0]], perm[A[l][1]])
        if rem < 130 or h + 99 > h_end: break
        if cm == 0:
            # the window's first lane is undecided: every lane is evaluated exactly (under the same states) and the round is
            # resolved again -- now it commits at least the first attempt
            rounds += 1
            ge = [False]*32; ac = [False]*32
            for l in lanes:
                i, j, bm = A[l]; si, sj = st[l]
                ge[l], ac[l] = evaluate(sA[l], i, j, si, sj)
            cm = resolve(ge, ac, [0]*32)
            assert cm & 1
            commit()
            P.update(mine=[False]*32, swaps=[False]*32, prom=[False]*32, adv=0, n=0)
            if rem < 130 or h + 99 > h_end: break
    done = nswap - rem
    log.sort()
    ok = (log == log_ref[:done])
    # the state after `done` attempts of the reference
    perm2 = list(range(K)); 
    for (s, si, sj, acc) in log_ref[:done]:
        pass
    msg = ('K=%d model=%s und=%.2f: attempts done %d/%d rounds %d (%.2f/round) log %s head %d vs ref next slot %d' % (
        K, model, p_und, done, nswap, rounds, done / max(rounds, 1), 'OK' if ok else 'MISMATCH', h, log_ref[done][0] if done < nswap else head_ref))
    assert h == (log_ref[done][0] if done < nswap else head_ref), msg
    assert ok, msg
    assert done >= nswap - 130


@pytest.mark.parametrize('K,model', [(256, 'ladder'), (256, 'flat'), (64, 'normal'), (16, 'ladder'), (4, 'flat'), (2, 'zeros'), (256, 'zeros')])
@pytest.mark.parametrize('p_und', [0.0, 0.03])
def test_walk2_round_logic_reproduces_reference(K, model, p_und):
    run(K, 4000, model, 7 + K, p_und)
