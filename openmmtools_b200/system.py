"""A light parameter record standing in for ``openmm.System`` on the replica-exchange path.

The reference builds OpenMM ``System`` objects (NonbondedForce / CustomNonbondedForce / CustomExternalForce,
/root/reference/openmmtools/testsystems.py:779-788,1956-1997; alchemy/alchemy.py:1539-2038) and OpenMM evaluates
them.  Here the same physics is described by plain numbers that the CUDA engine consumes; nothing in this module
imports OpenMM.
"""
import copy
import hashlib
import numpy as np

LJ, HARMONIC, MOLECULE = 'lj', 'harmonic', 'molecule'


class System:
    """Parameters of one of the supported systems (md units: nm, ps, dalton, kJ/mol)."""

    def __init__(self, kind, masses, box_vectors=None):
        if kind not in (LJ, HARMONIC, MOLECULE):
            raise ValueError('unsupported system kind %r' % (kind,))
        self.kind = kind
        self.masses = np.array(masses, dtype=np.float64)
        self.box_vectors = None if box_vectors is None else np.array(box_vectors, dtype=np.float64).reshape(3, 3)
        # -- Lennard-Jones fluid (NonbondedForce, CutoffPeriodic)
        self.sigma = None
        self.epsilon = None
        self.charge = None
        self.cutoff = None
        self.switching_distance = None
        self.use_switching_function = False
        self.use_dispersion_correction = False
        # -- alchemical modification (AbsoluteAlchemicalFactory)
        self.alchemical_atoms = None          # sorted tuple of atom indices, None = not alchemically modified
        self.annihilate_sterics = False
        self.softcore_alpha = 0.5
        self.softcore_a = 1.0
        self.softcore_b = 1.0
        self.softcore_c = 6.0
        self.alchemical_dispersion_correction = False
        # -- harmonic oscillator (CustomExternalForce global parameters)
        self.ho_K = None
        self.ho_x0 = (0.0, 0.0, 0.0)
        self.ho_U0 = 0.0
        # -- small molecule in vacuum (HarmonicBond/Angle, PeriodicTorsion, NonbondedForce NoCutoff + exceptions, HBonds
        #    constraints, CMMotionRemover): what AmberPrmtopFile.createSystem builds for AlanineDipeptideVacuum
        self.bonds = None          # float64[nb, 4]  i, j, K (kJ/mol/nm^2), r0 (nm)        1/2 K (r - r0)^2
        self.angles = None         # float64[na, 5]  i, j, k, K (kJ/mol/rad^2), theta0     1/2 K (t - t0)^2
        self.torsions = None       # float64[nt, 7]  i, j, k, l, n, phase, k (kJ/mol)      k (1 + cos(n phi - phase))
        self.exclusions = None     # int[ne, 2]      pairs without any nonbonded interaction (1-2, 1-3)
        self.exceptions = None     # float64[nx, 5]  i, j, q_i q_j (scaled), sigma, epsilon (scaled): the 1-4 pairs
        self.constraints = None    # float64[nc, 3]  i, j, distance (nm)
        self.remove_cm_motion = False
        # -- OpenMM-style global parameters with their defaults (read by GlobalParameterState.from_system)
        self.global_parameters = {}

    # -- OpenMM-flavoured accessors used by user code on this path
    def getNumParticles(self):
        return len(self.masses)

    def getParticleMass(self, i):
        from . import unit
        return self.masses[i] * unit.dalton

    def getDefaultPeriodicBoxVectors(self):
        from . import unit
        return unit.Quantity(self.box_vectors.copy(), unit.nanometer)

    def setDefaultPeriodicBoxVectors(self, a, b, c):
        from . import unit
        self.box_vectors = np.array([unit.to_md(a), unit.to_md(b), unit.to_md(c)], dtype=np.float64)

    def usesPeriodicBoundaryConditions(self):
        return self.kind == LJ

    def getNumConstraints(self):
        return 0 if self.constraints is None else len(self.constraints)

    @property
    def n_particles(self):
        return len(self.masses)

    @property
    def is_alchemical(self):
        return self.alchemical_atoms is not None

    def alchemical_mask(self):
        m = np.zeros(self.n_particles, np.uint8)
        if self.alchemical_atoms:
            m[list(self.alchemical_atoms)] = 1
        return m

    def copy(self):
        return copy.deepcopy(self)

    # -- the equivalent of ThermodynamicState._standard_system_hash (states.py:1447-1504): everything except the
    # thermodynamic parameters (temperature, lambdas, HO K/x0/U0 which are global parameters in the reference)
    def standard_hash(self):
        h = hashlib.sha1()
        h.update(self.kind.encode())
        h.update(self.masses.tobytes())
        if self.box_vectors is not None:
            h.update(self.box_vectors.tobytes())
        if self.kind == LJ:
            for a in (self.sigma, self.epsilon, self.charge):
                h.update(np.asarray(a, dtype=np.float64).tobytes())
            h.update(repr((self.cutoff, self.switching_distance, self.use_switching_function,
                           self.use_dispersion_correction, self.alchemical_atoms, self.annihilate_sterics,
                           self.softcore_alpha, self.softcore_a, self.softcore_b, self.softcore_c,
                           self.alchemical_dispersion_correction)).encode())
        if self.kind == MOLECULE:
            for a in (self.sigma, self.epsilon, self.charge, self.bonds, self.angles, self.torsions, self.exclusions,
                      self.exceptions, self.constraints):
                h.update(np.asarray(a, dtype=np.float64).tobytes())
            h.update(repr(self.remove_cm_motion).encode())
        return h.hexdigest()

    def __getstate__(self):
        return self.__dict__.copy()

    def __setstate__(self, d):
        self.__dict__.update(d)
