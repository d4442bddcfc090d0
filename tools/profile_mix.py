"""Run the swap-all mixing kernel alone on a synthetic high-acceptance energy matrix (for ncu / timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import numpy as np
from energy_models import energies
from openmmtools_b200._engine import Engine
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nswap = int(sys.argv[2]) if len(sys.argv) > 2 else K ** 3
model = sys.argv[3] if len(sys.argv) > 3 else 'flat'
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
e = Engine(0, K, K)
e.set_energies(energies(model, K, 99))
e.set_replica_states(np.arange(K))
e.mix_seed(1234)
for r in range(reps):
    e.phase_times(reset=True)
    t0 = time.time()
    st, nacc, nprop = e.mix_swap_all(nswap)
    dt = time.time() - t0
    ms = e.phase_times()['mix_ms']
    s = e.mix_stats()
    print('K=%d nswap=%d %s: %.2f ms (wall %.2f)  rounds %d  attempts/round %.2f  ns/round %.1f  acc %.3f  exact_exp %d' % (
        K, nswap, model, ms, dt * 1e3, s['rounds'], nswap / max(s['rounds'], 1), 1e6 * ms / max(s['rounds'], 1),
        nacc.sum() / nprop.sum(), s['exact_exp']))
