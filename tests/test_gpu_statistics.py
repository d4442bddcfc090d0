"""Ensemble-level validation with the reference's own thresholds: dynamics cannot be compared
bit for bit with OpenMM, so the sampled distributions are.

* harmonic-oscillator replica exchange: MBAR free-energy differences within 6 sigma of the analytic
  -1.5 ln(2 pi sigma_i^2) (/root/reference/openmmtools/tests/test_sampling.py:146,283-307), with a small
  self-contained MBAR solver (pymbar is not installed);
* kinetic temperature and <U> of the oscillator per state (tests/test_mcmc.py:97-248 style, 6 sigma);
* uniform state visitation under swap-all with zero energies (tests/test_mixing.py:76-92, chi-square).
"""
import numpy as np
import pytest
from openmmtools_b200 import unit, states, mcmc, testsystems, multistate
from helpers import KB

pytestmark = [pytest.mark.gpu]


def mbar(u_kn, N_k, tol=1e-10, maxit=10000):
    """Self-consistent MBAR: u_kn[k, n] reduced potential of sample n in state k; returns f_k (f_0 = 0)."""
    K, N = u_kn.shape
    f = np.zeros(K)
    logN = np.log(N_k)
    for _ in range(maxit):
        # log denominator per sample: logsumexp_k (log N_k + f_k - u_kn)
        a = logN[:, None] + f[:, None] - u_kn
        m = a.max(axis=0)
        logden = m + np.log(np.exp(a - m).sum(axis=0))
        b = -u_kn - logden[None, :]
        mb = b.max(axis=1)
        fnew = -(mb + np.log(np.exp(b - mb[:, None]).sum(axis=1)))
        fnew -= fnew[0]
        if np.abs(fnew - f).max() < tol:
            f = fnew
            break
        f = fnew
    return f


def test_harmonic_oscillator_free_energies_and_moments():
    """5 oscillators sigma_i = (1 + 0.2 i) A at 300 K (the reference's analytical test set-up, test_sampling.py:115-150),
    swap-all replica exchange with Langevin splitting dynamics."""
    T = 300.0
    kT = KB * T
    n = 5
    sigmas = np.array([(1.0 + 0.2 * i) * 0.1 for i in range(n)])     # nm
    Ks = kT / sigmas ** 2                                            # kJ/mol/nm^2
    tstates, sstates = [], []
    for i in range(n):
        ho = testsystems.HarmonicOscillator(K=Ks[i] * unit.kilojoule_per_mole / unit.nanometer ** 2, mass=12.0 * unit.amu)
        tstates.append(states.ThermodynamicState(ho.system, T * unit.kelvin))
        sstates.append(states.SamplerState(ho.positions))
    move = mcmc.LangevinSplittingDynamicsMove(timestep=2.0 * unit.femtosecond, collision_rate=20.0 / unit.picosecond,
                                              n_steps=100, reassign_velocities=False)
    s = multistate.ReplicaExchangeSampler(mcmc_moves=move, number_of_iterations=100000, seed=20240924)
    s.create(tstates, sstates)
    s.equilibrate(100)
    n_iter = 1500
    u_all = np.zeros((n_iter, n, n))
    st_all = np.zeros((n_iter, n), int)
    kin_all = np.zeros((n_iter, n))
    for it in range(n_iter):
        s.run(1)
        u_all[it] = s._energy_thermodynamic_states
        st_all[it] = s._replica_thermodynamic_states
        kin_all[it] = s._engine.get_replica_energies()[1]
    # subsample to decorrelate (100 steps x 2 fs with gamma = 20/ps: a few correlation times per iteration)
    sub = slice(0, n_iter, 3)
    u = u_all[sub]; st = st_all[sub]
    nsub = u.shape[0]
    u_kn = u.transpose(2, 0, 1).reshape(n, nsub * n)          # state l, samples (iteration, replica)
    N_k = np.array([(st == l).sum() for l in range(n)], float)
    f = mbar(u_kn, N_k)
    f_exact = -1.5 * np.log(2 * np.pi * sigmas ** 2)
    f_exact -= f_exact[0]
    # bootstrap error bars over iterations (blocks keep the replica structure)
    rng = np.random.default_rng(0)
    boots = []
    for _ in range(40):
        idx = rng.integers(0, nsub, nsub)
        ub = u[idx]; sb = st[idx]
        boots.append(mbar(ub.transpose(2, 0, 1).reshape(n, nsub * n), np.array([(sb == l).sum() for l in range(n)], float)))
    err = np.std(boots, axis=0) + 1e-12
    nsig = np.abs(f - f_exact)[1:] / err[1:]
    assert np.all(nsig < 6.0), (f, f_exact, err)
    assert np.all(err[1:] < 0.2)
    # moments per state: <U> = 1.5 kT, kinetic temperature = T (6 sigma with a correlation-corrected error)
    for l in range(n):
        mask = st == l
        ul = (u[..., l][mask]) * kT
        g = 3.0     # conservative statistical inefficiency after subsampling
        se = ul.std() / np.sqrt(len(ul) / g)
        assert abs(ul.mean() - 1.5 * kT) < 6 * se, (l, ul.mean(), 1.5 * kT, se)
        kl = kin_all[sub][mask]
        se_k = kl.std() / np.sqrt(len(kl) / g)
        assert abs(kl.mean() - 1.5 * kT) < 6 * se_k, (l, kl.mean(), 1.5 * kT, se_k)
    # replicas visit all states
    assert all(len(set(st_all[:, k])) == n for k in range(n))


def test_uniform_mixing_chi_square():
    """Reference tests/test_mixing.py:11-44,76-92: zero energies, 16 states, n_states**4 attempts per call; every replica
    must visit every state uniformly (chi-square p >= 0.001/16)."""
    from scipy import stats
    K, ncalls = 16, 1000
    hist = np.zeros((K, K))
    u = np.zeros((K, K))
    st = np.arange(K, dtype=np.int64)
    for c in range(ncalls):
        na = np.zeros((K, K), np.int64); npr = np.zeros((K, K), np.int64)
        multistate.ReplicaExchangeSampler._mix_all_replicas_numba(K ** 4, K, st, u, na, npr, seed=(4242 if c == 0 else None))
        hist[np.arange(K), st] += 1
    for k in range(K):
        chi2, p = stats.chisquare(hist[k])
        assert p >= 0.001 / K, (k, p)


def test_lj_fluid_kinetic_temperature_and_energy_conservation_limit():
    """Alchemical LJ fluid: kinetic temperature per replica equals the bath temperature; with gamma -> 0 and a small
    time step the total energy is conserved to float32 accuracy (force/energy consistency of the fast-math kernel)."""
    from test_gpu_sampler import lj_sampler
    s, asys, lambdas = lj_sampler(K=16, N=512, n_alch=10, n_steps=200, seed=3)
    s.run(12)
    kin = []
    for it in range(10):
        s.run(1)
        kin.append(s._engine.get_replica_energies()[1])
    T_kin = 2 * np.mean(kin, axis=0) / (3 * 512 * KB)
    assert np.all(np.abs(T_kin - 300.0) < 12.0), T_kin            # sigma_T ~ 300*sqrt(2/(3*512*10)) ~ 3.4 K
    # NVE limit
    from helpers import gpu_engine, lj_setup
    sset = lj_setup(N=256, n_alch=6, seed=21)
    e = gpu_engine(1, 2, 2, 256, box=(sset['L'],) * 3, r_cutoff=sset['rc'], r_switch=sset['rs'], use_switch=True)
    e.set_particles(sset['sigma'], sset['eps'], sset['mass'], sset['alch'])
    e.set_states([300.0, 300.0], [1.0, 0.5])
    e.set_positions(np.stack([sset['x']] * 2))
    e.set_replica_states(np.array([0, 1]))
    e.randomize_velocities(5)
    e.set_integrator(0.001, 0.0, 0, 'V R O R V')
    e.propagate(1, 0)
    p0, k0 = e.get_replica_energies()
    e.set_integrator(0.001, 0.0, 2000, 'V R O R V')
    e.propagate(1, 1)
    p1, k1 = e.get_replica_energies()
    drift = np.abs((p1 + k1) - (p0 + k0))
    assert np.all(drift < 0.05 * 256 * KB * 300 / 100), (drift, p0 + k0)   # << kT per atom
