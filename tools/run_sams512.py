"""The exchange machinery of BASELINE.json configs[4] at its stated size on a system this engine has: SAMSSampler over 512
lambda states of the alchemical Lennard-Jones fluid (512 atoms; the WaterBox itself needs PME + SETTLE, not built), K replicas
(default 8, one per GPU of the 8 x B200 the config names -- here all on one), 500 steps per iteration.  Times an iteration with
the host weight update (numpy on the fetched K x 512 matrix), with the device kernel per iteration (rx_sams_step) and in the
fused device loop (rx_sams_run_iterations), and checks that the three agree.
usage: run_sams512.py [iterations [K [M]]]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmmtools_b200 import testsystems, alchemy, states, mcmc, multistate, unit

n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 200
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
M = int(sys.argv[3]) if len(sys.argv) > 3 else 512
fluid = testsystems.LennardJonesFluid(nparticles=512)
asys = alchemy.AbsoluteAlchemicalFactory(disable_alchemical_dispersion_correction=True).create_alchemical_system(
    fluid.system, alchemy.AlchemicalRegion(alchemical_atoms=range(10), annihilate_sterics=False))
lam = [1.0 - l / (M - 1) for l in range(M)]
ts = states.create_thermodynamic_state_protocol(asys, {'lambda_sterics': lam, 'lambda_electrostatics': lam},
                                                constants={'temperature': 300.0 * unit.kelvin},
                                                composable_states=alchemy.AlchemicalState.from_system(asys))
ss = states.SamplerState(fluid.positions, box_vectors=asys.getDefaultPeriodicBoxVectors())


def make(device):
    move = mcmc.LangevinSplittingDynamicsMove(timestep=1.0 * unit.femtosecond, collision_rate=10.0 / unit.picosecond, n_steps=500)
    s = multistate.SAMSSampler(mcmc_moves=move, number_of_iterations=10 ** 9, seed=1234, gamma0=1.0, flatness_criteria='minimum-visits',
                               device_weight_update=device)
    s.create(ts, [ss] * K, storage=None)
    s.run(5)
    return s


out = {}
for name, device in (('host update', False), ('device kernel per iteration', True), ('fused device loop', True)):
    s = make(device)
    e = s._engine
    e.phase_times(reset=True)
    t0 = time.time()
    if name.startswith('fused'):
        s.run_fused(n_it)
    else:
        s.run(n_it)
    dt = time.time() - t0
    pt = e.phase_times()
    out[name] = (np.array(s._replica_thermodynamic_states), s._logZ.copy(), s._stage)
    print('%-28s %7.3f ms per iteration (%7.1f iterations/s); device phases: jump+update %.3f propagate %.3f energies %.3f ms' % (
        name, 1e3 * dt / n_it, n_it / dt, pt['mix_ms'] / n_it, pt['propagate_ms'] / n_it, pt['energies_ms'] / n_it))
a, b, c = out['host update'], out['device kernel per iteration'], out['fused device loop']
print('K=%d replicas, M=%d states, %d iterations: states equal (host, device, fused): %s %s   max |logZ_dev - logZ_host| %.2e   '
      'fused == per-iteration bitwise: %s   stage %d' % (K, M, n_it, np.array_equal(a[0], b[0]), np.array_equal(b[0], c[0]),
                                                        np.abs(a[1] - b[1]).max(), np.array_equal(b[1], c[1]), b[2]))
