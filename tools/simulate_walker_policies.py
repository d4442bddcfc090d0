"""CPU study (no GPU): how many rounds would the speculative walker need under different policies?

The exact swap-all chain of replicaexchange.py:321-349 is run sequentially on the MT19937 word stream (numba semantics, K a
power of two), recording for every 2-word slot whether an attempt starts there, its replicas and whether it changed the
permutation.  Rounds are then counted for:
  P0       the shipped walker: a round ends at the first visited slot that shares a replica with an earlier accepted,
           state-changing swap of the same round (window of W slots);
  P1       first-order forwarding: such a slot may stay in the round when exactly ONE earlier swap of the round touched it
           and that swap was itself evaluated normally (its new state could be forwarded by a shuffle);
  levels   the dependency depth of the true DAG (attempt waits for the last accepted swap that touched its replicas, and
           for its predecessor in the stream): the floor for any scheme without value prediction.
Usage: python tools/simulate_walker_policies.py [K] [nswap] [model]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import numpy as np
from energy_models import energies

K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nswap = int(sys.argv[2]) if len(sys.argv) > 2 else 300000
model = sys.argv[3] if len(sys.argv) > 3 else 'flat'
u = energies(model, K, 99)
words = np.random.RandomState(1234)._bit_generator.random_raw(4 * nswap + 64).astype(np.uint64)
mask = K - 1
perm = list(range(K))
ul = u.tolist()
slots = []   # (is_start, i, j, changed)
p = 0
import math
for t in range(nswap):
    i = int(words[p]) & mask; j = int(words[p + 1]) & mask
    si, sj = perm[i], perm[j]
    logp = -(ul[i][sj] + ul[j][si]) + ul[i][si] + ul[j][sj]
    p += 2
    if logp >= 0.0:
        acc = True; drew = False
    else:
        a, b = int(words[p]) >> 5, int(words[p + 1]) >> 6
        r = (a * 67108864.0 + b) / 9007199254740992.0
        acc = r < math.exp(logp); drew = True
        p += 2
    changed = acc and i != j
    if changed:
        perm[i], perm[j] = sj, si
    slots.append((True, i, j, changed))
    if drew:
        slots.append((False, -1, -1, False))
n_slots = len(slots)
print('K=%d model=%s attempts=%d slots=%d (%.2f slots/attempt)' % (K, model, nswap, n_slots, n_slots / nswap))


def rounds(W, forward):
    h = 0; n_rounds = 0; att = 0
    while h < n_slots:
        touched = {}          # replica -> number of accepted swaps of this round that touched it (and whether forwarded)
        s = h
        end = min(h + W, n_slots)
        while s < end:
            st, i, j, ch = slots[s]
            if st:
                ci, cj = touched.get(i), touched.get(j)
                stale = ci is not None or cj is not None
                fwd_lane = False
                if stale:
                    if not forward:
                        break
                    deps = [c for c in (ci, cj) if c is not None]
                    # forwardable: every touching swap was a normally evaluated one, and each replica was touched once
                    if any(c[0] > 1 or c[1] for c in deps) or s == h:
                        break
                    fwd_lane = True
                if ch:
                    for r in (i, j):
                        c = touched.get(r)
                        touched[r] = (1 if c is None else c[0] + 1, fwd_lane or (c is not None and c[1]))
                att += 1
            s += 1
        if s == h:   # the first lane itself can never be stale (fresh round): safety
            s = h + 1
        h = s
        n_rounds += 1
    return n_rounds


for W in (32, 64):
    r0 = rounds(W, False); r1 = rounds(W, True)
    print('window %2d slots: P0 %7d rounds (%.2f attempts/round)   P1 %7d rounds (%.2f attempts/round, x%.2f fewer)' % (
        W, r0, nswap / r0, r1, nswap / r1, r0 / r1))
# dependency levels
last = [0] * K
F = 0
for st, i, j, ch in slots:
    if not st:
        continue
    l = max(last[i], last[j]) + 1
    F = max(F, l)
    if ch:
        last[i] = F; last[j] = F
print('true dependency depth: %d levels (%.2f attempts/level)' % (F, nswap / F))
