"""SAMS on the device (csrc/rx_sams.cuh: rx_sams_set / rx_sams_step / rx_sams_run_iterations) against the host restatement of
/root/reference/openmmtools/multistate/sams.py:395-437, 477-501, 564-691 (multistate/sams.py, pinned bit for bit to golden
vectors lifted from the reference, tests/test_sams.py): the same jumps from the same MT19937 uniforms, logZ / weights / stage /
t0 / gamma to 1e-10, for every weight-update method, stage schedule and flatness criterion; the raw kernel against a numpy
restatement on a 512-state matrix (the size of BASELINE configs[4]); resume; the fused device loop."""
import numpy as np
import pytest
from scipy.special import logsumexp
from helpers import gpu_engine, KB
from openmmtools_b200 import testsystems, states, mcmc, multistate, unit, _lib

pytestmark = pytest.mark.gpu


def oscillator_ladder(n=5, T=300.0):
    kT = KB * T
    sigmas = np.array([(1.0 + 0.2 * i) * 0.1 for i in range(n)])
    tstates = []
    for i in range(n):
        ho = testsystems.HarmonicOscillator(K=(kT / sigmas[i] ** 2) * unit.kilojoule_per_mole / unit.nanometer ** 2, mass=12.0 * unit.amu)
        tstates.append(states.ThermodynamicState(ho.system, T * unit.kelvin))
    return tstates, ho, sigmas


def make_sampler(device, n_replicas=2, n_states=5, storage=None, **kw):
    tstates, ho, _ = oscillator_ladder(n_states)
    move = mcmc.LangevinSplittingDynamicsMove(timestep=2.0 * unit.femtosecond, collision_rate=20.0 / unit.picosecond, n_steps=50)
    s = multistate.SAMSSampler(mcmc_moves=move, number_of_iterations=10 ** 6, seed=4711, device_weight_update=device, **kw)
    s.create(tstates, [states.SamplerState(ho.positions)] * n_replicas, storage=storage)
    return s


@pytest.mark.parametrize('kw', [
    dict(weight_update_method='rao-blackwellized', flatness_criteria='minimum-visits'),
    dict(weight_update_method='optimal', flatness_criteria='minimum-visits'),
    dict(weight_update_method='rao-blackwellized', flatness_criteria='histogram-flatness', flatness_threshold=0.6),
    dict(weight_update_method='rao-blackwellized', flatness_criteria='logZ-flatness', flatness_threshold=0.05, gamma0=2.0),
    dict(weight_update_method='optimal', update_stages='one-stage'),
])
def test_device_update_follows_the_host_path(kw):
    host, dev = make_sampler(False, **kw), make_sampler(True, **kw)
    switched = False
    for it in range(120):
        host.run(1); dev.run(1)
        # propagation is the same kernel with the same noise in both samplers: identical energies as long as the jumps agree
        assert np.array_equal(host._replica_thermodynamic_states, dev._replica_thermodynamic_states), it
        assert np.array_equal(host._energy_thermodynamic_states, dev._energy_thermodynamic_states), it
        assert host._stage == dev._stage and host._t0 == dev._t0, (it, host._stage, dev._stage, host._t0, dev._t0)
        np.testing.assert_allclose(dev._logZ, host._logZ, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(dev.log_weights, host.log_weights, rtol=1e-10, atol=1e-12)
        assert np.array_equal(host._n_accepted_matrix, dev._n_accepted_matrix) and np.array_equal(host._n_proposed_matrix, dev._n_proposed_matrix)
        if host._last_gamma is not None:
            assert abs(host._last_gamma - dev._last_gamma) < 1e-13 * abs(host._last_gamma)
        switched |= host._stage == 1
    assert switched or kw.get('flatness_criteria') == 'logZ-flatness'
    # the host generator was advanced in lockstep with the device's copy of the stream
    assert np.array_equal(host._rng.get_state()[1], dev._rng.get_state()[1]) and host._rng.get_state()[2] == dev._rng.get_state()[2]
    assert np.array_equal(host._state_histogram, dev._state_histogram)
    assert np.array_equal(dev._engine.sams_get()['histogram'], dev._state_histogram)


def test_paths_can_be_switched_and_a_device_run_resumes(tmp_path):
    a = make_sampler(True, storage=multistate.MultiStateReporter(str(tmp_path / 'a'), checkpoint_interval=10),
                     flatness_criteria='minimum-visits')
    a.run(35)
    a.device_weight_update = False      # host path for a while: the device state is handed back and forth
    a.run(7)
    a.device_weight_update = True
    a.run(18)
    b = make_sampler(False, flatness_criteria='minimum-visits')
    b.run(60)
    assert np.array_equal(a._replica_thermodynamic_states, b._replica_thermodynamic_states)
    np.testing.assert_allclose(a._logZ, b._logZ, rtol=1e-10, atol=1e-12)
    del a
    c = multistate.SAMSSampler.from_storage(str(tmp_path / 'a'))
    assert c.device_weight_update and c.iteration == 60
    c.run(15); b.run(15)
    assert np.array_equal(c._replica_thermodynamic_states, b._replica_thermodynamic_states)
    np.testing.assert_allclose(c._logZ, b._logZ, rtol=1e-10, atol=1e-12)


def test_fused_device_loop_equals_iteration_by_iteration():
    a, b = make_sampler(True, flatness_criteria='minimum-visits'), make_sampler(True, flatness_criteria='minimum-visits')
    a.run(5); b.run(5)
    a.run(40)
    b.run_fused(40)
    assert b.iteration == a.iteration == 45
    assert np.array_equal(a._replica_thermodynamic_states, b._replica_thermodynamic_states)
    assert np.array_equal(a._logZ, b._logZ) and a._stage == b._stage and a._t0 == b._t0      # the same kernel: bitwise
    assert np.array_equal(a._state_histogram, b._state_histogram)
    assert np.array_equal(a._energy_thermodynamic_states, b._energy_thermodynamic_states)
    a.run(3); b.run(3)
    assert np.array_equal(a._replica_thermodynamic_states, b._replica_thermodynamic_states) and np.array_equal(a._logZ, b._logZ)


def numpy_sams_step(u, states_, log_w, logZ, log_pi, hist, stage, t0, iteration, rng, method, criteria, thr, gamma0, two_stage=True):
    """sams.py:395-437 restated once more on plain arrays (global jump; all replicas, then the update)."""
    K, M = u.shape
    logP = np.zeros((K, M))
    new = states_.copy()
    for r in range(K):
        lp = -u[r] + log_w
        lp -= logsumexp(lp)
        new[r] = rng.choice(np.arange(M), p=np.exp(lp))
        logP[r] = lp
    logZ = logZ.copy()
    gamma = None
    if iteration > 0:
        N = hist
        if two_stage and stage == 0 and N.sum() > 0:
            pi = np.exp(log_pi)
            adv = {0: np.all(N >= 1), 1: np.all(np.abs(pi - N / N.sum()) / pi < thr), 2: np.all(np.abs(logZ / gamma0) > thr)}[criteria]
            if adv or (t0 > 0 and iteration > t0):
                stage, t0 = 1, iteration - 1
        pi_star = np.exp(log_pi).min()
        t = float(iteration)
        gamma = gamma0 * min(pi_star, t ** -0.8) if stage == 0 else gamma0 * min(pi_star, 1.0 / (t - t0 + t0 ** 0.8))
        for r in range(K):
            if method == 0:
                logZ[new[r]] += gamma * np.exp(-log_pi[new[r]])
            else:
                logZ += gamma * np.exp(logP[r] - log_pi)
        if stage == 1:
            logZ -= logZ[0]
        log_w = log_pi - logZ
    return new, log_w, logZ, stage, t0, gamma


@pytest.mark.parametrize('K,M,method,criteria', [(1, 512, 1, 2), (3, 512, 0, 0), (4, 37, 1, 1), (2, 4096, 1, 0)])
def test_kernel_on_a_synthetic_matrix_at_the_size_of_config5(K, M, method, criteria):
    """rx_sams_step alone: K replicas over M states (BASELINE configs[4]: 512 lambda states), energy matrix set from the host,
    non-uniform targets, 30 iterations of a drifting matrix."""
    rng = np.random.default_rng(5)
    e = gpu_engine(_lib.RX_SYSTEM_HARMONIC, K, M, 1)
    e.set_particles(None, None, np.full(1, 12.0), None)
    e.set_states(np.full(M, 300.0), ho_K=np.linspace(100.0, 900.0, M), ho_x0=np.zeros((M, 3)))
    log_pi = np.log(rng.dirichlet(np.full(M, 20.0)))
    logZ = rng.normal(scale=0.3, size=M); logZ -= logZ[0]
    hist = np.zeros(M, np.int64)
    st = rng.integers(0, M, size=K).astype(np.int64)
    seed = 31337
    e.mix_seed(seed, _lib.RX_STREAM_NUMPY)
    e.set_replica_states(st)
    names_m = {0: 'optimal', 1: 'rao-blackwellized'}; names_c = {0: 'minimum-visits', 1: 'histogram-flatness', 2: 'logZ-flatness'}
    thr, gamma0 = (0.9 if criteria == 1 else 0.02), 1.5
    e.sams_set(log_pi, logZ, histogram=hist, gamma0=gamma0, flatness_threshold=thr, weight_update_method=names_m[method],
               flatness_criteria=names_c[criteria])
    host_rng = np.random.RandomState(seed)
    log_w = log_pi - logZ
    stage, t0 = 0, 0
    base = rng.normal(scale=3.0, size=(K, M)) + np.linspace(0.0, 6.0, M)[None, :]
    for it in range(0, 30):
        u = base + 0.3 * rng.normal(size=(K, M))
        e.set_energies(u)
        e.sams_step(it, update_weights=it > 0)
        r = e.sams_get()
        new, log_w, logZ, stage, t0, gamma = numpy_sams_step(u, st, log_w, logZ, log_pi, hist, stage, t0, it, host_rng, method, criteria,
                                                             thr, gamma0)
        assert np.array_equal(r['previous_states'], st) and np.array_equal(r['states'], new), it
        st = new
        np.add.at(hist, st, 1)
        assert np.array_equal(r['histogram'], hist)
        assert r['stage'] == stage and r['t0'] == t0
        np.testing.assert_allclose(r['logZ'], logZ, rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(r['log_weights'], log_w, rtol=1e-11, atol=1e-12)
        if gamma is not None:
            assert abs(r['gamma'] - gamma) <= 1e-13 * gamma
    assert e.mix_stream_position(_lib.RX_STREAM_NUMPY) == 2 * K * 30
    e.close()


# ---- the kernel against the golden vectors lifted from the reference itself (tests/golden/make_sams_golden.py runs
# /root/reference/openmmtools/multistate/sams.py's own _global_jump / _update_stage / _update_logZ_estimates) ------------------
from test_sams import G, energies_for


@pytest.mark.parametrize('tag', [str(t) for t in G['cases']])
def test_kernel_reproduces_the_reference_goldens(tag):
    _, K, M, s, stages, method, criteria = tag.split('_')
    K, M, seed = int(K[1:]), int(M[1:]), int(s[1:])
    e = gpu_engine(_lib.RX_SYSTEM_HARMONIC, K, M, 1)
    e.set_particles(None, None, np.full(1, 12.0), None)
    e.set_states(np.full(M, 300.0), ho_K=np.full(M, 100.0), ho_x0=np.zeros((M, 3)))
    e.mix_seed(seed, _lib.RX_STREAM_NUMPY)                  # the reference draws from numpy's global RandomState
    st0 = np.linspace(0, M - 1, K, dtype=int) if K > 1 else np.zeros(1, dtype=int)
    e.set_replica_states(st0.astype(np.int64))
    e.sams_set(np.zeros(M) - np.log(M), np.zeros(M), histogram=np.zeros(M, np.int64), gamma0=float(G[tag + '_gamma0']),
               weight_update_method=method, update_stages=stages, flatness_criteria=criteria,
               stage=1 if stages == 'one-stage' else 0, t0=0)
    prev = st0
    for it in range(1, 61):
        e.set_energies(energies_for(it, K, M, seed + 17))
        e.sams_step(it, update_weights=True)
        r = e.sams_get()
        i = it - 1
        assert np.array_equal(r['previous_states'], prev) and np.array_equal(r['states'], G[tag + '_states'][i]), it
        prev = r['states']
        assert r['stage'] == G[tag + '_stage'][i] and r['t0'] == G[tag + '_t0'][i], it
        np.testing.assert_allclose(r['logZ'], G[tag + '_logZ'][i], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(r['log_weights'], G[tag + '_log_weights'][i], rtol=1e-11, atol=1e-13)
        nacc = np.zeros((M, M), np.int64); nprop = np.zeros((M, M), np.int64)
        for c, n in zip(r['previous_states'], r['states']):
            nprop[c, :] += 1; nacc[c, n] += 1
        assert np.array_equal(nacc, G[tag + '_nacc'][i]) and np.array_equal(nprop, G[tag + '_nprop'][i])
    e.close()
