"""SAMSSampler: the state update restates the reference (sams.py:395-437,477-501,564-691) bit for bit against golden
vectors lifted from it (CPU, no engine needed); on the GPU a short expanded-ensemble run recovers the analytic free
energies of harmonic oscillators and survives a resume."""
import os
import numpy as np
import pytest
from openmmtools_b200 import multistate, states, unit, testsystems, mcmc
from energy_models import pseudo_normal
from helpers import KB

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'sams_golden.npz'))


def energies_for(it, K, M, key):
    x = pseudo_normal(K, key * 1000 + it)
    mu = 0.7 * np.arange(M, dtype=np.float64)
    return 0.5 * (3.0 * x[:, None] - mu[None, :]) ** 2 * 0.1 + 0.05 * mu[None, :]


@pytest.mark.parametrize('tag', [str(t) for t in G['cases']])
def test_state_update_matches_reference(tag):
    _, K, M, s, stages, method, criteria = tag.split('_')
    K, M, seed = int(K[1:]), int(M[1:]), int(s[1:])
    f = multistate.SAMSSampler(update_stages=stages, weight_update_method=method, flatness_criteria=criteria,
                               gamma0=float(G[tag + '_gamma0']))
    # drive the update logic without an engine: the attributes _pre_write_create would have set
    f._thermodynamic_states = [None] * M
    f._sampler_states = [None] * K
    f.locality = None
    f.log_target_probabilities = np.zeros(M) - np.log(M)
    f._logZ = np.zeros(M)
    f._initialize_stage()
    f._cached_state_histogram = np.zeros(M, dtype=int)
    f._replica_thermodynamic_states = np.linspace(0, M - 1, K, dtype=int) if K > 1 else np.zeros(1, dtype=int)
    f._n_accepted_matrix = np.zeros((M, M), np.int64); f._n_proposed_matrix = np.zeros((M, M), np.int64)
    f._neighborhoods = np.ones((K, M), np.int8)
    f._update_log_weights()
    f._rng = np.random.RandomState(seed)        # the reference draws from numpy's global RandomState
    for it in range(1, 61):
        f._iteration = it
        f._energy_thermodynamic_states = energies_for(it, K, M, seed + 17)
        f._mix_replicas()
        st, cnt = np.unique(f._replica_thermodynamic_states, return_counts=True)
        f._cached_state_histogram[st] += cnt
        i = it - 1
        assert np.array_equal(f._replica_thermodynamic_states, G[tag + '_states'][i]), it
        assert np.array_equal(f._logZ, G[tag + '_logZ'][i]), it
        assert np.array_equal(f.log_weights, G[tag + '_log_weights'][i]), it
        assert f._stage == G[tag + '_stage'][i] and f._t0 == G[tag + '_t0'][i], it
        assert np.array_equal(f._n_accepted_matrix, G[tag + '_nacc'][i]) and np.array_equal(f._n_proposed_matrix, G[tag + '_nprop'][i])


def test_option_validation_and_default_initial_states():
    with pytest.raises(ValueError):
        multistate.SAMSSampler(state_update_scheme='local-jump')      # the reference only allows global-jump (sams.py:246)
    with pytest.raises(ValueError):
        multistate.SAMSSampler(weight_update_method='bogus')
    d = multistate.MultiStateSampler._default_initial_thermodynamic_states
    assert list(d([0] * 5, [0] * 5)) == [0, 1, 2, 3, 4]
    assert list(d([0] * 5, [0])) == [0]
    assert list(d([0] * 5, [0] * 3)) == [0, 2, 4]
    assert list(d([0] * 3, [0] * 7)) == [0, 1, 2, 0, 1, 2, 0]


@pytest.mark.gpu
def test_sams_recovers_oscillator_free_energies_and_resumes(tmp_path):
    T = 300.0
    kT = KB * T
    n = 5
    sigmas = np.array([(1.0 + 0.2 * i) * 0.1 for i in range(n)])
    tstates = []
    for i in range(n):
        ho = testsystems.HarmonicOscillator(K=(kT / sigmas[i] ** 2) * unit.kilojoule_per_mole / unit.nanometer ** 2, mass=12.0 * unit.amu)
        tstates.append(states.ThermodynamicState(ho.system, T * unit.kelvin))
    sstates = [states.SamplerState(ho.positions)] * 2          # 2 replicas over 5 states
    move = mcmc.LangevinSplittingDynamicsMove(timestep=2.0 * unit.femtosecond, collision_rate=20.0 / unit.picosecond, n_steps=100)

    def make(path):
        s = multistate.SAMSSampler(mcmc_moves=move, number_of_iterations=3000, seed=99, gamma0=1.0, flatness_criteria='minimum-visits')
        s.create(tstates, sstates, storage=multistate.MultiStateReporter(path, checkpoint_interval=50))
        return s

    a = make(str(tmp_path / 'a'))
    assert list(a._replica_thermodynamic_states) == [0, 4]
    a.run(1500)
    f_exact = -1.5 * np.log(2 * np.pi * sigmas ** 2); f_exact -= f_exact[0]
    f_est = -(a._logZ - a._logZ[0])
    assert a._stage == 1
    assert np.abs(f_est - f_exact).max() < 0.35, (f_est, f_exact)
    hist = a._state_histogram / a._state_histogram.sum()
    assert np.all(np.abs(hist - 0.2) < 0.08), hist
    # resume: continue 100 more iterations in two ways
    a.run(100)
    b = make(str(tmp_path / 'b'))
    b.run(1550)
    del b
    c = multistate.SAMSSampler.from_storage(str(tmp_path / 'b'))
    assert c.iteration == 1550
    c.run(50)
    assert np.array_equal(c._replica_thermodynamic_states, a._replica_thermodynamic_states)
    assert np.array_equal(c._logZ, a._logZ)
