"""Parity ON the BASELINE.json configurations themselves (configs[1]: K = M = 64, configs[2]: K = M = 256; 512 atoms,
500 BAOAB steps of 1 fs, swap-all, seed 1234), through the public API:
  * the device energy matrix equals the oracle's on the positions the device holds (1e-5 relative, the north-star bar);
  * the permutation and both count matrices equal the oracle's replay of the reference loop on the same MT19937 stream
    and the same matrix, bit for bit;
  * a full-length (500-step) launch with the default dual neighbour list equals the all-pairs launch, and the first steps of
    a 512-atom launch equal the oracle integrator fed with the same noise across at least one re-partition of the list."""
import numpy as np
import pytest
from openmmtools_b200 import unit, states, alchemy, mcmc, testsystems, multistate, _backend
from helpers import KB, lj_setup, oracle_system, gpu_engine, device_noise

pytestmark = pytest.mark.gpu


def baseline_sampler(K):
    fluid = testsystems.LennardJonesFluid(nparticles=512)
    factory = alchemy.AbsoluteAlchemicalFactory(disable_alchemical_dispersion_correction=True)
    asys = factory.create_alchemical_system(fluid.system, alchemy.AlchemicalRegion(alchemical_atoms=range(10), annihilate_sterics=False))
    lambdas = [1.0 - l / (K - 1) for l in range(K)]
    tstates = states.create_thermodynamic_state_protocol(
        asys, {'lambda_sterics': lambdas, 'lambda_electrostatics': lambdas}, constants={'temperature': 300.0 * unit.kelvin},
        composable_states=alchemy.AlchemicalState.from_system(asys))
    sstate = states.SamplerState(fluid.positions, box_vectors=asys.getDefaultPeriodicBoxVectors())
    move = mcmc.LangevinSplittingDynamicsMove(timestep=1.0 * unit.femtosecond, collision_rate=10.0 / unit.picosecond,
                                              n_steps=500, reassign_velocities=False, splitting='V R O R V')
    s = multistate.ReplicaExchangeSampler(mcmc_moves=move, number_of_iterations=100, replica_mixing_scheme='swap-all', seed=1234)
    s.create(tstates, [sstate])
    return s, asys, lambdas


def oracle_matrix(asys, lambdas, x):
    from oracle import oracle
    L = asys.box_vectors[0, 0]
    osys = oracle.LJSystem(asys.sigma, asys.epsilon, asys.masses, asys.alchemical_mask(), (L, L, L), asys.cutoff,
                           asys.switching_distance, use_switch=True)
    K = len(lambdas)
    off = np.full(K, _backend.lj_dispersion_correction(asys))
    return osys.energy_matrix(x, np.array(lambdas), np.full(K, 1.0 / (KB * 300.0)), off)


@pytest.mark.parametrize('K,iterations', [(64, 2), (256, 1)])
def test_baseline_config_iterations_match_oracle(K, iterations):
    from oracle import oracle
    s, asys, lambdas = baseline_sampler(K)
    mt = oracle.MT(1234 & 0xFFFFFFFF)
    perm = np.arange(K, dtype=np.int64)
    s._compute_energies()
    u_prev = s._energy_thermodynamic_states.copy()
    x0 = np.stack([st.positions.value_in_unit(unit.nanometer) for st in s.sampler_states])
    ref0 = oracle_matrix(asys, lambdas, x0)
    assert np.abs(u_prev - ref0).max() / np.abs(ref0).max() < 1e-5
    for it in range(1, iterations + 1):
        s.run(1)
        na = np.zeros((K, K), np.int64); npr = np.zeros((K, K), np.int64)
        oracle.mix_swap_all(mt, K ** 3, perm, u_prev, na, npr)
        assert np.array_equal(perm, s._replica_thermodynamic_states), it
        assert np.array_equal(na, s._n_accepted_matrix) and np.array_equal(npr, s._n_proposed_matrix), it
        assert npr.sum() == 2 * K ** 3
        x = np.stack([st.positions.value_in_unit(unit.nanometer) for st in s.sampler_states])
        assert not np.array_equal(x, x0)          # the replicas did move
        ref = oracle_matrix(asys, lambdas, x)
        u_prev = s._energy_thermodynamic_states.copy()
        rel = np.abs(u_prev - ref).max() / np.abs(ref).max()
        assert rel < 1e-5, (it, rel)


def _engine(s, K, lambdas, temps, dt, gamma, n_steps):
    e = gpu_engine(1, K, K, s['N'], box=(s['L'],) * 3, r_cutoff=s['rc'], r_switch=s['rs'], use_switch=True)
    e.set_particles(s['sigma'], s['eps'], s['mass'], s['alch'])
    e.set_states(temps, lambdas)
    e.set_integrator(dt, gamma, n_steps, 'V R O R V')
    return e


def test_full_length_default_list_launch_equals_all_pairs(monkeypatch):
    """N = 512, 300 K, 1 fs, 10/ps, 500 steps, default skins (0.05 / 0.40 nm): the dual-list launch against the all-pairs
    launch of the same kernel -- a pair missing from a list shows up as a different trajectory."""
    N, K = 512, 4
    s = lj_setup(N=N, n_alch=10, seed=41)
    lambdas = np.array([1.0, 0.7, 0.3, 0.0]); temps = np.full(K, 300.0)
    rng = np.random.default_rng(6)
    v0 = rng.normal(scale=np.sqrt(KB * 300.0 / s['mass'][0]), size=(K, N, 3)).astype(np.float32).astype(np.float64)
    out = []
    for no_list in (False, True):
        if no_list:
            monkeypatch.setenv('RX_NO_VERLET', '1')
        e = _engine(s, K, lambdas, temps, 0.001, 10.0, 500)
        e.set_positions(np.stack([s['x']] * K)); e.set_velocities(v0); e.set_replica_states(np.arange(K))
        e.propagate(4321, 1)
        out.append((e.get_positions(), e.get_replica_energies()[0]))
        e.close()
    (xa, pa), (xb, pb) = out
    d = xa - xb
    d -= s['L'] * np.round(d / s['L'])
    assert np.median(np.abs(d)) < 2e-5, np.median(np.abs(d))     # float32 summation order only
    assert np.abs(d).max() < 5e-3, np.abs(d).max()
    assert np.abs(pa - pb).max() < 0.02 * max(1.0, np.abs(pb).max())


def test_512_atoms_same_noise_steps_cross_a_repartition():
    """20 steps at 2 fs and 600 K on 512 atoms move some atom by more than half the inner skin (0.025 nm), i.e. the
    inner list is re-partitioned inside the launch; positions, velocities and energies against the oracle integrator fed
    with the device's noise."""
    N, K, n_steps = 512, 2, 20
    s = lj_setup(N=N, n_alch=10, seed=43)
    lambdas = np.array([1.0, 0.4]); temps = np.array([600.0, 600.0])
    dt, gamma = 0.002, 1.0
    rng = np.random.default_rng(8)
    v0 = rng.normal(scale=np.sqrt(KB * 600.0 / s['mass'][0]), size=(K, N, 3)).astype(np.float32).astype(np.float64)
    e = _engine(s, K, lambdas, temps, dt, gamma, n_steps)
    x0 = np.stack([s['x']] * K)
    e.set_positions(x0); e.set_velocities(v0); e.set_replica_states(np.arange(K))
    seed, iteration = 0xABCDEF12345, 3
    e.propagate(seed, iteration)
    xg, vg = e.get_positions(), e.get_velocities()
    pot, kin = e.get_replica_energies()
    osys = oracle_system(s)
    moved = 0.0
    for k in range(K):
        x = x0[k].copy(); v = v0[k].copy()
        noise = device_noise(seed, iteration, k, N, n_steps)
        U = osys.langevin(x, v, noise, lambdas[k], KB * temps[k], dt, gamma, n_steps, 'VRORV')
        moved = max(moved, np.sqrt(((x - x0[k]) ** 2).sum(axis=1)).max())
        xw = x - s['L'] * np.floor(x / s['L'])
        d = xg[k] - xw
        d -= s['L'] * np.round(d / s['L'])
        assert np.abs(d).max() < 1e-4, (k, np.abs(d).max())
        assert np.abs(vg[k] - v).max() < 1e-3, (k, np.abs(vg[k] - v).max())
        assert abs(pot[k] - U) < 1e-3 * max(1.0, abs(U)), (pot[k], U)
    assert moved > 0.025, moved     # the launch did cross a re-partition
    e.close()
