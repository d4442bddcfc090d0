"""End-to-end ReplicaExchangeSampler runs on the GPU through the public (reference-shaped) API."""
import numpy as np
import pytest
from openmmtools_b200 import unit, states, alchemy, mcmc, testsystems, multistate, _backend
from helpers import KB

pytestmark = pytest.mark.gpu


def lj_sampler(K=16, N=128, n_alch=4, n_steps=25, scheme='swap-all', seed=1234, **kw):
    fluid = testsystems.LennardJonesFluid(nparticles=N)
    asys = alchemy.AbsoluteAlchemicalFactory(disable_alchemical_dispersion_correction=True).create_alchemical_system(
        fluid.system, alchemy.AlchemicalRegion(alchemical_atoms=range(n_alch)))
    lambdas = [1.0 - l / (K - 1) for l in range(K)]
    tstates = states.create_thermodynamic_state_protocol(
        asys, {'lambda_sterics': lambdas}, constants={'temperature': 300.0 * unit.kelvin},
        composable_states=alchemy.AlchemicalState.from_system(asys))
    sstate = states.SamplerState(fluid.positions, box_vectors=asys.getDefaultPeriodicBoxVectors())
    move = mcmc.LangevinSplittingDynamicsMove(timestep=1.0 * unit.femtosecond, n_steps=n_steps)
    s = multistate.ReplicaExchangeSampler(mcmc_moves=move, number_of_iterations=1000, replica_mixing_scheme=scheme,
                                          seed=seed, **kw)
    s.create(tstates, [sstate])
    return s, asys, lambdas


def oracle_energy(asys, lambdas, x):
    from oracle import oracle
    L = asys.box_vectors[0, 0]
    osys = oracle.LJSystem(asys.sigma, asys.epsilon, asys.masses, asys.alchemical_mask(), (L, L, L), asys.cutoff,
                           asys.switching_distance, use_switch=True)
    K = len(lambdas)
    off = np.full(K, _backend.lj_dispersion_correction(asys))
    return osys.energy_matrix(x, np.array(lambdas), np.full(K, 1.0 / (KB * 300.0)), off)


@pytest.mark.parametrize('scheme', ['swap-all', 'swap-neighbors'])
def test_iteration_history_matches_oracle_replay(scheme):
    """Each iteration: the device energy matrix equals the oracle's on the device's positions (1e-5 relative), and the
    permutation + swap statistics equal the oracle's mixing of that matrix on the same MT19937 stream (bit-exact)."""
    from oracle import oracle
    K = 16
    x0 = np.stack([np.asarray(testsystems.LennardJonesFluid(nparticles=128).positions.value_in_unit(unit.nanometer), np.float64)] * K)
    s2, asys, lambdas = lj_sampler(K=K, scheme=scheme, seed=78)
    seed = 78 & 0xFFFFFFFF if scheme == 'swap-all' else (78 >> 16) & 0xFFFFFFFF
    mt = oracle.MT(seed)
    perm = np.arange(K, dtype=np.int64)
    s2._compute_energies()
    u_prev = s2._energy_thermodynamic_states.copy()
    ref0 = oracle_energy(asys, lambdas, x0)
    assert np.abs(u_prev - ref0).max() / np.abs(ref0).max() < 1e-5
    for it in range(1, 5):
        s2.run(1)
        na = np.zeros((K, K), np.int64); npr = np.zeros((K, K), np.int64)
        if scheme == 'swap-all':
            oracle.mix_swap_all(mt, K ** 3, perm, u_prev, na, npr)
        else:
            oracle.mix_swap_neighbors(mt, perm, u_prev, na, npr)
        assert np.array_equal(perm, s2._replica_thermodynamic_states), it
        assert np.array_equal(na, s2._n_accepted_matrix) and np.array_equal(npr, s2._n_proposed_matrix)
        x = np.stack([st.positions.value_in_unit(unit.nanometer) for st in s2.sampler_states])
        ref = oracle_energy(asys, lambdas, x)
        u_prev = s2._energy_thermodynamic_states.copy()
        assert np.abs(u_prev - ref).max() / np.abs(ref).max() < 1e-5
        assert s2.iteration == it


def test_fused_loop_equals_phase_by_phase():
    """rx_run_iterations (no host round trips) and the Python-level run() produce identical trajectories."""
    a, _, _ = lj_sampler(K=16, seed=5)
    b, _, _ = lj_sampler(K=16, seed=5)
    a.run(3)
    b._compute_energies()
    b._engine.run_iterations(3, 'swap-all', b._seed, 1)
    assert np.array_equal(a._engine.get_replica_states(), b._engine.get_replica_states())
    assert np.array_equal(a._engine.get_energies(), b._engine.get_energies())
    assert np.array_equal(a._engine.get_positions(), b._engine.get_positions())


def test_host_resident_states_round_trip_is_equivalent():
    a, _, _ = lj_sampler(K=8, seed=9)
    b, _, _ = lj_sampler(K=8, seed=9, host_resident_states=True)
    a.run(3); b.run(3)
    assert np.array_equal(a._replica_thermodynamic_states, b._replica_thermodynamic_states)
    xa = np.stack([s._positions for s in a.sampler_states]); xb = np.stack([s._positions for s in b.sampler_states])
    assert np.array_equal(xa, xb)
    assert a.sampler_states[0].potential_energy is not None and a.sampler_states[0].kinetic_energy is not None


def test_config1_harmonic_oscillator_three_temperatures():
    """BASELINE.json configs[0]: HarmonicOscillator, 3 temperature states, 10 iterations."""
    ho = testsystems.HarmonicOscillator()
    temps = [300.0, 310.0, 320.0]
    tstates = [states.ThermodynamicState(ho.system, T * unit.kelvin) for T in temps]
    sstate = states.SamplerState(ho.positions)
    move = mcmc.LangevinSplittingDynamicsMove(timestep=1.0 * unit.femtosecond, n_steps=500)
    hist = []
    for rep in range(2):
        s = multistate.ReplicaExchangeSampler(mcmc_moves=move, number_of_iterations=10, seed=1234)
        s.create(tstates, [sstate])
        h = []
        for it in range(10):
            s.run(1)
            h.append(s._replica_thermodynamic_states.copy())
            assert s._n_proposed_matrix.sum() == 2 * 27
        hist.append(np.array(h))
        assert s.is_completed and s.iteration == 10
        # u[k,l] = beta_l * K/2 |x_k|^2
        x = np.stack([st._positions for st in s.sampler_states])
        ref = np.array([[0.5 * ho.system.ho_K * (x[k] ** 2).sum() / (KB * T) for T in temps] for k in range(3)])
        assert np.allclose(s._energy_thermodynamic_states, ref, rtol=1e-5)
    assert np.array_equal(hist[0], hist[1])          # reproducible for a fixed seed


def test_reference_static_mixing_entry_point():
    """tests/test_mixing.py of the reference calls ReplicaExchangeSampler._mix_all_replicas_numba directly."""
    from oracle import oracle
    K = 16
    u = np.zeros((K, K))
    st = np.arange(K, dtype=np.int64); na = np.zeros((K, K), np.int64); npr = np.zeros((K, K), np.int64)
    multistate.ReplicaExchangeSampler._mix_all_replicas_numba(K ** 4, K, st, u, na, npr, seed=1234)
    st_o = np.arange(K, dtype=np.int64); nao = np.zeros((K, K), np.int64); npo = np.zeros((K, K), np.int64)
    oracle.mix_swap_all(oracle.MT(1234), K ** 4, st_o, u, nao, npo)
    assert np.array_equal(st, st_o) and np.array_equal(na, nao) and np.array_equal(npr, npo)


def test_thermodynamic_state_reduced_potential_and_single_move():
    fluid = testsystems.LennardJonesFluid(nparticles=128)
    ts = states.ThermodynamicState(fluid.system, 300 * unit.kelvin)
    ss = states.SamplerState(fluid.positions, box_vectors=fluid.system.getDefaultPeriodicBoxVectors())
    u0 = ts.reduced_potential(ss)
    from oracle import oracle
    s = fluid.system; L = s.box_vectors[0, 0]
    osys = oracle.LJSystem(s.sigma, s.epsilon, s.masses, s.alchemical_mask(), (L, L, L), s.cutoff, s.switching_distance)
    U, _ = osys.energy(ss._positions, 1.0)
    ref = (U + _backend.lj_dispersion_correction(s)) / (KB * 300)
    assert u0 == pytest.approx(ref, rel=1e-5)
    move = mcmc.LangevinSplittingDynamicsMove(n_steps=50)
    move.apply(ts, ss)
    assert ss.velocities is not None and ss.potential_energy is not None and not ss.has_nan()
    assert ts.reduced_potential(ss) != u0


def test_parallel_tempering_on_lj_fluid():
    """ParallelTemperingSampler (paralleltempering.py:109-173): temperatures follow the reference's np.logspace spacing,
    u[k,l] = beta_l U(x_k), swaps reproduce the oracle's mixing of that matrix."""
    from oracle import oracle
    import math
    fluid = testsystems.LennardJonesFluid(nparticles=128)
    ts = states.ThermodynamicState(fluid.system, 300 * unit.kelvin)
    ss = states.SamplerState(fluid.positions, box_vectors=fluid.system.getDefaultPeriodicBoxVectors())
    move = mcmc.LangevinSplittingDynamicsMove(timestep=1.0 * unit.femtosecond, n_steps=50)
    s = multistate.ParallelTemperingSampler(mcmc_moves=move, number_of_iterations=100, seed=11)
    K = 8
    s.create(ts, ss, min_temperature=300 * unit.kelvin, max_temperature=600 * unit.kelvin, n_temperatures=K)
    T = [st.temperature.value_in_unit(unit.kelvin) for st in s._thermodynamic_states]
    ref_T = list(np.logspace(np.log10(300.0), np.log10(600.0), num=K))
    assert np.allclose(T, ref_T)
    s.run(3)
    u = s._energy_thermodynamic_states
    # rows are one potential energy scaled by beta_l
    kT = KB * np.array(T)
    U = u * kT[None, :]
    assert np.allclose(U, U[:, :1], rtol=1e-9)
    mt = oracle.MT(11)
    perm = np.arange(K, dtype=np.int64)
    s2 = multistate.ParallelTemperingSampler(mcmc_moves=move, number_of_iterations=100, seed=11)
    s2.create(ts, ss, temperatures=[t * unit.kelvin for t in ref_T])
    s2._compute_energies()
    u_prev = s2._energy_thermodynamic_states.copy()
    for it in range(3):
        s2.run(1)
        na = np.zeros((K, K), np.int64); npr = np.zeros((K, K), np.int64)
        oracle.mix_swap_all(mt, K ** 3, perm, u_prev, na, npr)
        assert np.array_equal(perm, s2._replica_thermodynamic_states)
        u_prev = s2._energy_thermodynamic_states.copy()
    assert np.array_equal(s2._replica_thermodynamic_states, s._replica_thermodynamic_states)
    with pytest.raises(ValueError):
        multistate.ParallelTemperingSampler().create(ts, ss, temperatures=[300 * unit.kelvin], n_temperatures=3)


def test_unsampled_states_energy_matrix():
    """Reference tests/test_sampling.py:1668-1717: the energy matrices (sampled and unsampled states) equal an independent
    double loop through reduced_potential."""
    K = 6
    fluid = testsystems.LennardJonesFluid(nparticles=128)
    asys = alchemy.AbsoluteAlchemicalFactory().create_alchemical_system(fluid.system, alchemy.AlchemicalRegion(alchemical_atoms=range(4)))
    lam = [1.0 - l / (K - 1) for l in range(K)]
    a = alchemy.AlchemicalState.from_system(asys)
    tstates = states.create_thermodynamic_state_protocol(asys, {'lambda_sterics': lam}, constants={'temperature': 300 * unit.kelvin},
                                                         composable_states=a)
    un = states.create_thermodynamic_state_protocol(asys, {'lambda_sterics': [0.37, 1.0], 'temperature': [280, 350] * unit.kelvin},
                                                    composable_states=a)
    sstate = states.SamplerState(fluid.positions, box_vectors=asys.getDefaultPeriodicBoxVectors())
    s = multistate.ReplicaExchangeSampler(mcmc_moves=mcmc.LangevinSplittingDynamicsMove(n_steps=30), number_of_iterations=2, seed=4)
    s.create(tstates, [sstate], unsampled_thermodynamic_states=un)
    s.run()
    assert s._energy_unsampled_states.shape == (K, 2)
    ss = s.sampler_states
    for k in (0, K - 1):
        for l in (0, 2, K - 1):
            assert s._energy_thermodynamic_states[k, l] == pytest.approx(tstates[l].reduced_potential(ss[k]), rel=1e-9)
        for l in range(2):
            assert s._energy_unsampled_states[k, l] == pytest.approx(un[l].reduced_potential(ss[k]), rel=1e-9)


def test_locality_neighborhoods():
    """locality (multistatesampler.py:1263-1281): only the band around each replica's state is (re)written; the
    reference requires swap-neighbors with it (replicaexchange.py:228-230)."""
    with pytest.raises(ValueError):
        multistate.ReplicaExchangeSampler(locality=1)                 # default swap-all is rejected, as in the reference
    s, asys, lambdas = lj_sampler(K=16, scheme='swap-neighbors', seed=6, locality=2)
    s.run(4)
    nb = s._neighborhoods
    for k, st in enumerate(s._replica_thermodynamic_states):
        band = np.zeros(16, np.int8); band[max(0, st - 2):min(16, st + 3)] = 1
        assert np.array_equal(nb[k], band)
    full = s._engine.get_energies()
    assert np.array_equal(s._energy_thermodynamic_states[nb == 1], full[nb == 1])
    # same seed without locality: identical permutation history (swap-neighbors only reads entries inside the band)
    t, _, _ = lj_sampler(K=16, scheme='swap-neighbors', seed=6)
    t.run(4)
    assert np.array_equal(s._replica_thermodynamic_states, t._replica_thermodynamic_states)


def test_minimize_lowers_energy_and_meets_force_tolerance():
    """MultiStateSampler.minimize (multistatesampler.py:611-647): every replica is relaxed in its own state until the
    RMS force component is below the tolerance; checked with the oracle's f64 forces on the returned positions."""
    from oracle import oracle
    K = 4
    s, asys, lambdas = lj_sampler(K=K, N=128, n_steps=5)
    with pytest.raises(TypeError):
        s.minimize(tolerance=1.0 * unit.nanometer)
    L = asys.box_vectors[0, 0]
    osys = oracle.LJSystem(asys.sigma, asys.epsilon, asys.masses, asys.alchemical_mask(), (L, L, L), asys.cutoff,
                           asys.switching_distance, use_switch=True)
    x0 = np.stack([st.positions.value_in_unit(unit.nanometer) for st in s.sampler_states])
    v0 = [st.velocities for st in s.sampler_states]
    tol = 5.0
    s.minimize(tolerance=tol * unit.kilojoules_per_mole / unit.nanometers)
    x1 = np.stack([st.positions.value_in_unit(unit.nanometer) for st in s.sampler_states])
    assert np.all(np.isfinite(x1)) and np.all(x1 >= 0) and np.all(x1 <= L)
    for k in range(K):
        e0 = osys.energy(x0[k], lambdas[k])[0]
        e1, _, f1 = osys.energy(x1[k], lambdas[k], forces=True)
        assert e1 < e0, (k, e0, e1)
        rms = np.sqrt(np.mean(np.asarray(f1) ** 2))
        assert rms < tol * 1.02, (k, rms)
        assert s.sampler_states[k].potential_energy is None  # new positions invalidate the cached energy
    info = s._last_minimization
    assert np.all(info['rms_force'] <= tol) and np.all(info['iterations'] > 0)
    for a, b in zip(v0, [st.velocities for st in s.sampler_states]):
        assert (a is None and b is None) or np.array_equal(a.value_in_unit(a.unit), b.value_in_unit(b.unit))
    # a capped descent stops at max_iterations
    s.minimize(tolerance=1e-6 * unit.kilojoules_per_mole / unit.nanometers, max_iterations=7)
    assert np.all(s._last_minimization['iterations'] == 7)
    s.run(1)  # the sampler continues from the minimized configuration
    assert s.iteration == 1


def test_minimize_harmonic_oscillator_reaches_the_well():
    ho = testsystems.HarmonicOscillator()
    x = np.array([[0.3, -0.2, 0.1]])
    hs = multistate.ReplicaExchangeSampler(mcmc_moves=mcmc.LangevinSplittingDynamicsMove(n_steps=5), number_of_iterations=2, seed=3)
    hs.create([states.ThermodynamicState(ho.system, T * unit.kelvin) for T in (300, 320)],
              [states.SamplerState(unit.Quantity(x, unit.nanometer))])
    hs.minimize(tolerance=0.5 * unit.kilojoules_per_mole / unit.nanometers)
    for st in hs.sampler_states:
        r = np.linalg.norm(st.positions.value_in_unit(unit.nanometer))
        assert r < 0.3 * 0.01, r


def test_graft_entry_smoke_runs():
    """The driver's smoke(): one small iteration through the public API, checked against the oracle inside."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as g
    g.smoke()


def test_nan_restart_retries_only_the_failed_replica_from_the_iteration_start():
    """Restart policy of mcmc.py:706-759: a replica whose propagation went NaN is retried from the state it had when the
    iteration began (device-side snapshot), the others keep their result (they are not propagated twice, and nobody is
    rewound to an older host copy).  Replica 2 is given absurd velocities after two clean iterations, so that every
    attempt fails and the run ends with SimulationNaNError; the healthy replicas must equal those of a clean run."""
    from openmmtools_b200.multistate.utils import SimulationNaNError
    K = 8
    clean, _, _ = lj_sampler(K=K, seed=21, scheme=None)
    sick, _, _ = lj_sampler(K=K, seed=21, scheme=None)
    clean.run(3)
    sick.run(2)
    v = sick._engine.get_velocities()
    v[2] = 1e30
    sick._engine.set_velocities(v[2:3], first=2)
    with pytest.raises(SimulationNaNError):
        sick.run(1)
    xs, xc = sick._engine.get_positions(), clean._engine.get_positions()
    for k in range(K):
        if k != 2:
            assert np.array_equal(xs[k], xc[k]), k        # propagated exactly once, from the right state


@pytest.mark.gpu
def test_sampler_accepts_one_move_per_state():
    """mcmc_moves as a list with different moves (multistatesampler.py:906-910): the sampler hands them to the engine state
    by state; the replicas in the hot, short-timestep states run their own move."""
    from openmmtools_b200 import testsystems, states, mcmc, multistate, unit
    ho = testsystems.HarmonicOscillator()
    ts = [states.ThermodynamicState(ho.system, T * unit.kelvin) for T in (300, 400, 500)]
    moves = [mcmc.LangevinSplittingDynamicsMove(timestep=2.0 * unit.femtosecond, n_steps=30),
             mcmc.LangevinSplittingDynamicsMove(timestep=1.0 * unit.femtosecond, n_steps=60),
             mcmc.LangevinSplittingDynamicsMove(timestep=0.5 * unit.femtosecond, n_steps=120, splitting='O V R V O')]
    s = multistate.ReplicaExchangeSampler(mcmc_moves=moves, number_of_iterations=5, seed=3)
    s.create(ts, [states.SamplerState(ho.positions)], storage=None)
    s.run()
    assert s.iteration == 5 and np.all(np.isfinite(s._energy_thermodynamic_states))
    assert [m.n_steps for m in s.mcmc_moves] == [30, 60, 120]
