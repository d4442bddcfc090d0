"""Where the end-to-end iteration (ReplicaExchangeSampler.run with host-resident sampler states) spends its host time:
cProfile over n iterations of the bench workload.  usage: profile_e2e.py [mixing [n [K]]]"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from openmmtools_b200 import mcmc, multistate, unit

mixing = sys.argv[1] if len(sys.argv) > 1 else 'swap-neighbors'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
fluid, asys, tstates, sstate, lambdas = bench.build_workload(K, 512)
move = mcmc.LangevinSplittingDynamicsMove(timestep=1.0 * unit.femtosecond, collision_rate=10.0 / unit.picosecond, n_steps=500,
                                          splitting='V R O R V', reassign_velocities=False)
s = multistate.ReplicaExchangeSampler(mcmc_moves=move, number_of_iterations=10 ** 9, replica_mixing_scheme=mixing, seed=1234)
s.create(tstates, [sstate], storage=None)
s.host_resident_states = True
s.run(3)
t0 = time.time(); s.run(n); dt = time.time() - t0
print('%s K=%d: %.3f ms per end-to-end iteration' % (mixing, K, 1e3 * dt / n))
pr = cProfile.Profile(); pr.enable(); s.run(n); pr.disable()
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(18)
e = s._engine
pt = e.phase_times()
print('device phases (ms, cumulative):', pt)
