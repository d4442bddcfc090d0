"""Quick phase timing of the alchemical LJ configuration straight through the engine (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmmtools_b200 import testsystems, unit
from openmmtools_b200._engine import Engine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 500
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 1
fl = testsystems.LennardJonesFluid(nparticles=512)
s = fl.system
L = s.box_vectors[0, 0]
e = Engine(1, K, K, 512, box=(L, L, L), r_cutoff=s.cutoff, r_switch=s.switching_distance, use_switch=True)
alch = np.zeros(512, np.uint8); alch[:10] = 1
e.set_particles(s.sigma, s.epsilon, s.masses, alch)
lam = 1.0 - np.arange(K) / (K - 1)
e.set_states(np.full(K, 300.0), lam)
e.set_integrator(0.001, 10.0, n_steps, 'V R O R V')
x = np.asarray(fl.positions.value_in_unit(unit.nanometer), np.float64)
e.set_positions(np.stack([x] * K))
e.set_replica_states(np.arange(K))
e.randomize_velocities(2024)
e.mix_seed(1234, 0)
u = e.compute_energies()
print('u[0,:4]', u[0, :4], 'spread', u.max() - u.min())
for it in range(iters):
    e.phase_times(reset=True)
    t0 = time.time()
    e.run_iterations(batch, 'swap-all', 7, it * batch)
    dt = time.time() - t0
    pt = e.phase_times()
    nacc, nprop = e.get_mix_counts()
    print('iter %d wall %.1f ms  mix %.2f prop %.2f energy %.2f  acc %.3f  words %d' % (
        it, dt * 1e3 / batch, pt['mix_ms'] / batch, pt['propagate_ms'] / batch, pt['energies_ms'] / batch, nacc.sum() / max(nprop.sum(), 1),
        e.mix_stream_position(0)), e.mix_stats())
pot, kin = e.get_replica_energies()
print('T_kin', (2 * kin / (3 * 512 * 8.31446261815324e-3))[:4], 'pot', pot[:3])
