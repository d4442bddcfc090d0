"""MultiStateReporter container (CPU) and checkpoint/resume through it (GPU): reference semantics of
/root/reference/openmmtools/multistate/multistatereporter.py:763-999,1184-1201 and multistatesampler.py:956-1047."""
import os
import numpy as np
import pytest
from openmmtools_b200 import multistate, states, unit, testsystems
from openmmtools_b200.multistate import MultiStateReporter


def test_reporter_records_and_commit_marker(tmp_path):
    path = str(tmp_path / 'run.store')
    r = MultiStateReporter(path, open_mode='w', checkpoint_interval=2)
    K, M, N = 3, 3, 5
    r.set_dimensions(K, M, N)
    r.write_dict('options', {'number_of_iterations': 7, 'replica_mixing_scheme': 'swap-all'})
    rng = np.random.default_rng(0)
    hist = []
    for it in range(5):
        u = rng.normal(size=(K, M)); st = rng.permutation(K); na = rng.integers(0, 9, (M, M)); npr = na + 1
        r.write_energies(u, np.ones((K, M), np.int8), np.zeros((K, 0)), it)
        r.write_replica_thermodynamic_states(st, it)
        r.write_mixing_statistics(na, npr, it)
        ss = [states.SamplerState(rng.normal(size=(N, 3)), velocities=rng.normal(size=(N, 3)), box_vectors=np.eye(3) * 2.0)
              for _ in range(K)]
        wrote = r.write_sampler_states(ss, it, extra={'seed': 5, 'iteration': it})
        assert wrote == (it % 2 == 0)
        if it < 4:
            r.write_last_iteration(it)       # iteration 4 is written but NOT committed (crash before the marker)
        hist.append((u, st, na, npr, ss))
    r.close()
    r2 = MultiStateReporter(path, open_mode='r')
    assert r2.checkpoint_interval == 2
    assert r2.read_last_iteration(last_checkpoint=False) == 3
    assert r2.read_last_iteration() == 2                       # last committed checkpoint
    assert r2.read_checkpoint_iterations() == [0, 2, 4]
    e, nb, un = r2.read_energies()
    assert e.shape == (4, K, M) and nb.dtype == np.int8        # the uncommitted record is not exposed
    for it in range(4):
        assert np.array_equal(r2.read_energies(it)[0], hist[it][0])
        assert np.array_equal(r2.read_replica_thermodynamic_states(it), hist[it][1])
        a, p = r2.read_mixing_statistics(it)
        assert np.array_equal(a, hist[it][2]) and np.array_equal(p, hist[it][3]) and a.dtype == np.int32
    ss = r2.read_sampler_states(2)
    assert np.array_equal(ss[1]._positions, hist[2][4][1]._positions)
    assert np.array_equal(ss[1]._velocities, hist[2][4][1]._velocities)
    assert r2.read_sampler_states(1) is None
    assert r2.read_checkpoint_extra(2) == {'seed': 5, 'iteration': 2}
    assert r2.read_dict('options')['replica_mixing_scheme'] == 'swap-all'
    with pytest.raises(ValueError):
        MultiStateReporter(path).open('x')


def test_analysis_particle_streams_follow_the_reference_intervals(tmp_path):
    """analysis_particle_indices / position_interval / velocity_interval (multistatereporter.py:106-116, 1686-1692):
    the listed particles are stored at their own intervals, as float32, next to the sparse checkpoints."""
    path = str(tmp_path / 'an.store')
    r = MultiStateReporter(path, open_mode='w', checkpoint_interval=10, analysis_particle_indices=(4, 1), position_interval=2,
                           velocity_interval=0)
    assert r.analysis_particle_indices == (4, 1) and r.position_interval == 2 and r.velocity_interval == 0
    K, M, N = 2, 2, 6
    r.set_dimensions(K, M, N)
    rng = np.random.default_rng(1)
    kept = {}
    for it in range(6):
        ss = [states.SamplerState(rng.normal(size=(N, 3)), velocities=rng.normal(size=(N, 3)), box_vectors=np.eye(3) * 3.0)
              for _ in range(K)]
        assert r.wants_analysis_states(it) == (it % 2 == 0)
        r.write_sampler_states(ss, it)
        r.write_last_iteration(it)
        kept[it] = ss
    r.close()
    r2 = MultiStateReporter(path, open_mode='r')
    assert r2.analysis_particle_indices == (4, 1) and r2.position_interval == 2 and r2.velocity_interval == 0
    assert r2.read_sampler_states(3, analysis_particles_only=True) is None
    got = r2.read_sampler_states(4, analysis_particles_only=True)
    assert len(got) == K and got[0].n_particles == 2 and got[0].velocities is None
    want = kept[4][1]._positions[[4, 1]].astype(np.float32).astype(np.float64)
    assert np.array_equal(got[1]._positions, want)
    assert np.array_equal(got[1]._box_vectors, np.eye(3) * 3.0)
    assert r2.read_checkpoint_iterations() == [0]                   # full states only at the checkpoint interval
    no = MultiStateReporter(str(tmp_path / 'none.store'), open_mode='w')
    assert no.wants_analysis_states(0) is False and no.position_interval == 1 and no.velocity_interval == 1


@pytest.mark.gpu
def test_resume_is_bit_identical_to_uninterrupted_run(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from test_gpu_sampler import lj_sampler

    def make(path, interval):
        s, asys, lambdas = lj_sampler(K=16, seed=2024)
        return s, asys, lambdas

    # uninterrupted: 6 iterations with storage
    from openmmtools_b200 import alchemy, mcmc
    fluid = testsystems.LennardJonesFluid(nparticles=128)
    asys = alchemy.AbsoluteAlchemicalFactory(disable_alchemical_dispersion_correction=True).create_alchemical_system(
        fluid.system, alchemy.AlchemicalRegion(alchemical_atoms=range(4)))
    K = 16
    lambdas = [1.0 - l / (K - 1) for l in range(K)]
    tstates = states.create_thermodynamic_state_protocol(asys, {'lambda_sterics': lambdas}, constants={'temperature': 300.0 * unit.kelvin},
                                                         composable_states=alchemy.AlchemicalState.from_system(asys))
    sstate = states.SamplerState(fluid.positions, box_vectors=asys.getDefaultPeriodicBoxVectors())
    move = mcmc.LangevinSplittingDynamicsMove(timestep=1.0 * unit.femtosecond, n_steps=25)

    def new_sampler(path):
        s = multistate.ReplicaExchangeSampler(mcmc_moves=move, number_of_iterations=6, seed=2024)
        s.create(tstates, [sstate], storage=MultiStateReporter(path, checkpoint_interval=2))
        return s

    a = new_sampler(str(tmp_path / 'a.store'))
    a.run()
    assert a.iteration == 6 and a.is_completed
    ua = a._energy_thermodynamic_states.copy(); pa = a._replica_thermodynamic_states.copy()
    xa = np.stack([s._positions for s in a.sampler_states])
    ra = MultiStateReporter(str(tmp_path / 'a.store'), open_mode='r')
    assert ra.read_last_iteration(last_checkpoint=False) == 6
    assert np.array_equal(ra.read_energies(6)[0], ua) and np.array_equal(ra.read_replica_thermodynamic_states(6), pa)
    assert multistate.ReplicaExchangeSampler.read_status(str(tmp_path / 'a.store')).is_completed

    # interrupted after 5 iterations (last checkpoint = 4), resumed from storage
    b = new_sampler(str(tmp_path / 'b.store'))
    b.run(5)
    del b
    st = multistate.ReplicaExchangeSampler.read_status(str(tmp_path / 'b.store'))
    assert st.iteration == 5 and not st.is_completed
    c = multistate.ReplicaExchangeSampler.from_storage(str(tmp_path / 'b.store'))
    assert c.iteration == 4 and c.replica_mixing_scheme == 'swap-all' and c.number_of_iterations == 6
    c.run()
    assert c.iteration == 6
    assert np.array_equal(c._replica_thermodynamic_states, pa)
    assert np.array_equal(c._energy_thermodynamic_states, ua)
    assert np.array_equal(np.stack([s._positions for s in c.sampler_states]), xa)
    # the online free-energy estimate (multistatesampler.py:1625-1664) is stored per iteration and carried over the resume
    assert a._last_mbar_f_k.shape == (K,) and a._last_mbar_f_k[0] == 0.0 and np.all(np.isfinite(a._last_mbar_f_k))
    assert np.array_equal(c._last_mbar_f_k, a._last_mbar_f_k)
    assert np.array_equal(ra.read_online_analysis_data(6, 'f_k')['f_k'], a._last_mbar_f_k)
    assert np.array_equal(ra.read_online_analysis_data(None, 'f_k')['f_k'], a._last_mbar_f_k)
    fe = ra.read_online_analysis_data(6, 'free_energy')['free_energy']
    assert fe[0] == a._last_mbar_f_k[-1] and np.isinf(fe[1])
    rb = MultiStateReporter(str(tmp_path / 'b.store'), open_mode='r')
    for it in range(7):
        assert np.array_equal(rb.read_energies(it)[0], ra.read_energies(it)[0]), it
        assert np.array_equal(rb.read_mixing_statistics(it)[0], ra.read_mixing_statistics(it)[0]), it
