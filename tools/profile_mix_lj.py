"""Mixing kernel on the REAL alchemical-LJ energy matrix (after a few propagation iterations); dumps the matrix."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmmtools_b200 import testsystems, unit
from openmmtools_b200._engine import Engine
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fl = testsystems.LennardJonesFluid(nparticles=512)
s = fl.system; L = s.box_vectors[0, 0]
e = Engine(1, K, K, 512, box=(L, L, L), r_cutoff=s.cutoff, r_switch=s.switching_distance, use_switch=True)
alch = np.zeros(512, np.uint8); alch[:10] = 1
e.set_particles(s.sigma, s.epsilon, s.masses, alch)
e.set_states(np.full(K, 300.0), 1.0 - np.arange(K) / (K - 1))
e.set_integrator(0.001, 10.0, 500, 'V R O R V')
x = np.asarray(fl.positions.value_in_unit(unit.nanometer), np.float64)
e.set_positions(np.stack([x] * K)); e.set_replica_states(np.arange(K)); e.randomize_velocities(2024); e.mix_seed(1234, 0)
e.compute_energies()
e.run_iterations(3, 'swap-all', 7, 0)
u = e.get_energies()
np.save('gpurun_out/u_lj_%d.npy' % K, u)
print('u range per row: min %.3g max %.3g; global [%.3f, %.3f]' % (np.ptp(u, axis=1).min(), np.ptp(u, axis=1).max(), u.min(), u.max()))
d = np.abs(np.diff(u, axis=1))
print('adjacent-state |du|: min %.3g median %.3g; exact zeros %d' % (d.min(), np.median(d), (d == 0).sum()))
for rep in range(2):
    e.phase_times(reset=True)
    e.mix_swap_all(K ** 3, fetch=False)
    print('mix %.1f ms' % e.phase_times()['mix_ms'], e.mix_stats())
