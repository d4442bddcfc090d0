"""ctypes wrapper over oracle/liborc.so (the CPU restatement in rx_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU arms.
The product package (openmmtools_b200/) must never import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, 'liborc.so')


def build(force=False):
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ('rx_oracle.c', 'rx_oracle_mol.c'))
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < newest:
        subprocess.check_call(['make', '-C', _HERE, '-B', 'liborc.so'], stdout=subprocess.DEVNULL)
    return _LIB


class LJParams(C.Structure):
    _fields_ = [('n_atoms', C.c_int32), ('use_switch', C.c_int32), ('annihilate_sterics', C.c_int32),
                ('pad', C.c_int32), ('box', C.c_double * 3), ('r_cutoff', C.c_double), ('r_switch', C.c_double),
                ('alpha', C.c_double), ('a', C.c_double), ('b', C.c_double), ('c', C.c_double)]


class DynParams(C.Structure):
    _fields_ = [('kind', C.c_int32), ('n_steps', C.c_int32), ('dt', C.c_double), ('gamma', C.c_double),
                ('kT', C.c_double), ('lam', C.c_double), ('ho_K', C.c_double), ('ho_x0', C.c_double * 3),
                ('ho_U0', C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_rand.restype = C.c_double
        _lib.orc_randint.restype = C.c_int64
        _lib.orc_randint.argtypes = [C.c_void_p, C.c_int64]
        _lib.orc_mt_next32.restype = C.c_uint32
        _lib.orc_lj_energy.restype = C.c_double
        _lib.orc_lj_dispersion_correction.restype = C.c_double
        _lib.orc_ho_energy.restype = C.c_double
        _lib.orc_mix_swap_all.restype = C.c_int64
        _lib.orc_mix_swap_neighbors.restype = C.c_int64
        _lib.orc_mol_energy.restype = C.c_double
        _lib.orc_mol_langevin.restype = C.c_double
        _lib.orc_mol_kinetic.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class MT:
    """MT19937 state with numba / numpy-legacy semantics."""

    def __init__(self, seed):
        self.buf = C.create_string_buffer(lib().orc_mt_sizeof())
        lib().orc_mt_seed(self.buf, C.c_uint32(seed & 0xFFFFFFFF))

    def randint(self, n):
        return int(lib().orc_randint(self.buf, C.c_int64(n)))

    def rand(self):
        return float(lib().orc_rand(self.buf))

    def next32(self):
        return int(lib().orc_mt_next32(self.buf))


def mix_swap_all(mt, nswap, states, u, n_acc, n_prop):
    K = len(states)
    assert states.dtype == np.int64 and u.dtype == np.float64 and u.flags.c_contiguous
    lib().orc_mix_swap_all(mt.buf, C.c_int64(nswap), C.c_int64(K), _p(states), _p(u), _p(n_acc), _p(n_prop))


def mix_swap_neighbors(mt, states, u, n_acc, n_prop):
    K = len(states)
    r = lib().orc_mix_swap_neighbors(mt.buf, C.c_int64(K), _p(states), _p(u), _p(n_acc), _p(n_prop))
    assert r == 0


class LJSystem:
    """Alchemical LJ fluid parameters in md units (float64 arrays)."""

    def __init__(self, sigma, epsilon, mass, alch_mask, box, r_cutoff, r_switch, use_switch=True,
                 alpha=0.5, a=1.0, b=1.0, c=6.0, annihilate_sterics=False):
        self.sigma = np.ascontiguousarray(sigma, np.float64)
        self.eps = np.ascontiguousarray(epsilon, np.float64)
        self.mass = np.ascontiguousarray(mass, np.float64)
        self.alch = np.ascontiguousarray(alch_mask, np.uint8)
        self.N = len(self.sigma)
        self.p = LJParams(self.N, int(use_switch), int(annihilate_sterics), 0, (C.c_double * 3)(*box),
                          r_cutoff, r_switch, alpha, a, b, c)

    def energy(self, xyz, lam, forces=False):
        xyz = np.ascontiguousarray(xyz, np.float64)
        comp = np.zeros(3)
        f = np.zeros((self.N, 3)) if forces else None
        U = lib().orc_lj_energy(C.byref(self.p), _p(self.sigma), _p(self.eps), _p(self.alch), _p(xyz),
                                C.c_double(lam), _p(comp), _p(f) if forces else None)
        return (U, comp, f) if forces else (U, comp)

    def dispersion_correction(self):
        eps_nb = np.where(self.alch != 0, 0.0, self.eps)   # alchemy.py:1909
        return lib().orc_lj_dispersion_correction(C.byref(self.p), _p(self.sigma), _p(eps_nb))

    def energy_row(self, xyz, lambdas, betas, offsets=None):
        xyz = np.ascontiguousarray(xyz, np.float64)
        lambdas = np.ascontiguousarray(lambdas, np.float64); betas = np.ascontiguousarray(betas, np.float64)
        out = np.zeros(len(lambdas))
        off = None if offsets is None else np.ascontiguousarray(offsets, np.float64)
        lib().orc_lj_energy_row(C.byref(self.p), _p(self.sigma), _p(self.eps), _p(self.alch), _p(xyz),
                                C.c_int(len(lambdas)), _p(lambdas), _p(betas), _p(off) if off is not None else None,
                                _p(out))
        return out

    def energy_matrix(self, x, lambdas, betas, offsets=None, threads=0):
        x = np.ascontiguousarray(x, np.float64)
        count = x.shape[0]
        lambdas = np.ascontiguousarray(lambdas, np.float64); betas = np.ascontiguousarray(betas, np.float64)
        off = None if offsets is None else np.ascontiguousarray(offsets, np.float64)
        u = np.zeros((count, len(lambdas)))
        lib().orc_energy_matrix(C.byref(self.p), _p(self.sigma), _p(self.eps), _p(self.alch), C.c_int(count), _p(x),
                                C.c_int(len(lambdas)), _p(lambdas), _p(betas), _p(off) if off is not None else None,
                                _p(u), C.c_int(threads))
        return u

    def langevin(self, x, v, noise, lam, kT, dt, gamma, n_steps, program='VRORV'):
        """In-place V/R/O splitting steps with injected noise [n_steps*nO, N, 3]."""
        d = DynParams(0, n_steps, dt, gamma, kT, lam, 0.0, (C.c_double * 3)(0, 0, 0), 0.0)
        U = C.c_double()
        noise = np.ascontiguousarray(noise, np.float64)
        r = lib().orc_langevin_steps(C.byref(self.p), C.byref(d), program.replace(' ', '').encode(), _p(self.sigma),
                                     _p(self.eps), _p(self.alch), _p(self.mass), _p(x), _p(v), _p(noise), C.byref(U))
        assert r == 0, r
        return U.value

    def propagate_replicas(self, x, v, lambdas, kTs, dt, gamma, n_steps, seed, threads=0):
        count = x.shape[0]
        lambdas = np.ascontiguousarray(lambdas, np.float64); kTs = np.ascontiguousarray(kTs, np.float64)
        lib().orc_propagate_replicas(C.byref(self.p), _p(self.sigma), _p(self.eps), _p(self.alch), _p(self.mass),
                                     C.c_int(count), _p(x), _p(v), _p(lambdas), _p(kTs), C.c_double(dt),
                                     C.c_double(gamma), C.c_int(n_steps), C.c_uint64(seed), C.c_int(threads))


def ho_energy(xyz, K, x0, U0, forces=False):
    xyz = np.ascontiguousarray(xyz, np.float64)
    N = xyz.shape[0]
    x0 = np.ascontiguousarray(x0, np.float64)
    f = np.zeros((N, 3)) if forces else None
    U = lib().orc_ho_energy(C.c_int(N), _p(xyz), C.c_double(K), _p(x0), C.c_double(U0), _p(f) if forces else None)
    return (U, f) if forces else U


def ho_langevin(x, v, mass, noise, K, x0, U0, kT, dt, gamma, n_steps, program='VRORV'):
    N = x.shape[0]
    p = LJParams(N, 0, 0, 0, (C.c_double * 3)(1e3, 1e3, 1e3), 1.0, 0.5, 0.5, 1, 1, 6)
    d = DynParams(1, n_steps, dt, gamma, kT, 0.0, K, (C.c_double * 3)(*x0), U0)
    U = C.c_double()
    mass = np.ascontiguousarray(mass, np.float64)
    noise = np.ascontiguousarray(noise, np.float64)
    r = lib().orc_langevin_steps(C.byref(p), C.byref(d), program.replace(' ', '').encode(), None, None, None,
                                 _p(mass), _p(x), _p(v), _p(noise), C.byref(U))
    assert r == 0, r
    return U.value


def max_threads():
    return int(lib().orc_max_threads())


class _OrcMol(C.Structure):
    _fields_ = [('n_atoms', C.c_int), ('n_bonds', C.c_int), ('n_angles', C.c_int), ('n_torsions', C.c_int), ('n_excl', C.c_int),
                ('n_exc', C.c_int), ('n_cons', C.c_int), ('remove_cm', C.c_int),
                ('mass', C.c_void_p), ('charge', C.c_void_p), ('sigma', C.c_void_p), ('eps', C.c_void_p),
                ('bonds', C.c_void_p), ('angles', C.c_void_p), ('torsions', C.c_void_p), ('excl', C.c_void_p),
                ('exc', C.c_void_p), ('cons', C.c_void_p)]


class Molecule:
    """A small molecule in vacuum (rx_oracle_mol.c): bonds/angles/torsions, all-pairs Coulomb + LJ with exclusions and 1-4
    exceptions, distance constraints.  `system` is an openmmtools_b200.system.System of kind 'molecule' (plain arrays)."""

    def __init__(self, system):
        f = lambda a, w: np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1, w))
        self.mass = np.ascontiguousarray(system.masses, np.float64)
        self.charge = np.ascontiguousarray(system.charge, np.float64)
        self.sigma = np.ascontiguousarray(system.sigma, np.float64)
        self.eps = np.ascontiguousarray(system.epsilon, np.float64)
        self.bonds, self.angles, self.torsions = f(system.bonds, 4), f(system.angles, 5), f(system.torsions, 7)
        self.excl = np.ascontiguousarray(np.asarray(system.exclusions, np.int64).reshape(-1, 2))
        self.exc = f(system.exceptions, 5)
        self.cons = f(system.constraints if system.constraints is not None else np.zeros((0, 3)), 3)
        self.N = len(self.mass)
        self.s = _OrcMol(self.N, len(self.bonds), len(self.angles), len(self.torsions), len(self.excl), len(self.exc),
                         len(self.cons), int(bool(system.remove_cm_motion)), _p(self.mass).value, _p(self.charge).value,
                         _p(self.sigma).value, _p(self.eps).value, _p(self.bonds).value, _p(self.angles).value,
                         _p(self.torsions).value, _p(self.excl).value, _p(self.exc).value, _p(self.cons).value)

    def energy(self, x, forces=False):
        x = np.ascontiguousarray(x, np.float64)
        f = np.zeros_like(x) if forces else None
        U = lib().orc_mol_energy(C.byref(self.s), _p(x), _p(f) if forces else None)
        return (U, f) if forces else U

    def kinetic(self, v):
        return lib().orc_mol_kinetic(C.byref(self.s), _p(np.ascontiguousarray(v, np.float64)))

    def langevin(self, x, v, noise, kT, dt, gamma, n_steps, program='VRORV', tol=1e-10):
        """In-place constrained V/R/O splitting steps with injected noise [n_steps*nO, N, 3]; returns the final potential."""
        assert x.dtype == np.float64 and v.dtype == np.float64 and x.flags.c_contiguous and v.flags.c_contiguous
        noise = np.ascontiguousarray(noise, np.float64)
        return lib().orc_mol_langevin(C.byref(self.s), _p(x), _p(v), _p(noise), C.c_double(kT), C.c_double(dt), C.c_double(gamma),
                                      C.c_int(n_steps), program.replace(' ', '').encode(), C.c_double(tol))
