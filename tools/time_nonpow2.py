"""Swap-all at K not a power of two against the reference loop on one host core
(oracle/rx_oracle.c: the same loop numba compiles).  usage: time_nonpow2.py [model] [K ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import numpy as np
from energy_models import energies
from openmmtools_b200._engine import Engine
from oracle import oracle
model = 'ladder'   # low acceptance (2 %): nearly every attempt draws a uniform; 'flat': a fine alchemical ladder
args = sys.argv[1:]
if args and not args[0].isdigit(): model = args.pop(0)
for K in [int(a) for a in args] or [96, 100, 200]:
    u = energies(model, K, 5)
    e = Engine(0, K, K, 0)
    e.set_energies(u); e.set_replica_states(np.arange(K)); e.mix_seed(77, 0)
    n = K ** 3
    e.mix_swap_all(n)                       # warm-up (stream allocation)
    e.phase_times(reset=True)
    t0 = time.time(); st, na, npr = e.mix_swap_all(n); t_gpu = time.time() - t0
    mt = oracle.MT(77); perm = np.arange(K, dtype=np.int64)
    a = np.zeros((K, K), np.int64); b = np.zeros((K, K), np.int64)
    oracle.mix_swap_all(mt, n, perm, u, a, b)
    t0 = time.time(); oracle.mix_swap_all(mt, n, perm, u, a, b); t_cpu = time.time() - t0
    ms = e.mix_stats()
    print(model + ' K=%d: %d attempts  device %.1f ms (%.1f ns/attempt; walker %.1f ms, %d rounds, %.2f attempts/round, %.0f ns/round)   '
          'host core %.1f ms (%.1f ns/attempt)   same result: %s'
          % (K, n, 1e3 * t_gpu, 1e9 * t_gpu / n, ms['walker_ms'], ms['rounds'], n / max(ms['rounds'], 1), 1e6 * ms['walker_ms'] / max(ms['rounds'], 1),
             1e3 * t_cpu, 1e9 * t_cpu / n, np.array_equal(st, perm)))
    e.close()
