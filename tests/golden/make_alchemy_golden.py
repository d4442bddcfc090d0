#!/usr/bin/env python
"""Golden data for the alchemical force decomposition, from the REAL reference code (build container only).

``AbsoluteAlchemicalFactory`` (class body lifted by AST from /root/reference/openmmtools/alchemy/alchemy.py, together
with ``AlchemicalRegion``) runs its ``_alchemically_modify_NonbondedForce`` on recording stand-ins for
``openmm.NonbondedForce`` / ``CustomNonbondedForce`` / ``CustomBondForce`` (OpenMM itself is not installable here).
The forces it builds -- energy expressions, per-particle parameters, interaction groups, global parameters, cutoff /
switch / long-range-correction flags -- are (1) recorded and (2) evaluated with numpy on a small periodic LJ
configuration for several lambda values (NonbondedForce: Lorentz-Berthelot LJ inside the cutoff; CustomNonbondedForce:
the Lepton expression over the interaction groups), with the switching function off so that only what the reference
source defines enters the number.

Output: tests/golden/alchemy_golden.npz
"""
import ast, collections, copy, itertools, json, logging, os, re, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, HERE)
from openmmtools_b200 import unit as u
from helpers import lj_setup
from make_integrator_golden import eval_lepton_program


# ----------------------------------------------------------------------------------------------- recording stand-ins
class _Force:
    def __init__(self):
        self.globals = collections.OrderedDict()
        self.force_group = 0

    def addGlobalParameter(self, name, value): self.globals[name] = float(u.to_md(value)); return len(self.globals) - 1
    def setForceGroup(self, g): self.force_group = g
    def getForceGroup(self): return self.force_group


class NonbondedForce(_Force):
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)

    def __init__(self):
        super().__init__()
        self.particles, self.exceptions, self.offsets = [], [], []
        self.method, self.cutoff, self.use_switch, self.switch_distance = self.NoCutoff, 1.0 * u.nanometer, False, -1.0 * u.nanometer
        self.dispersion, self.rf_dielectric, self.ewald_tol = True, 78.3, 5e-4

    def addParticle(self, q, s, e): self.particles.append([q, s, e]); return len(self.particles) - 1
    def getNumParticles(self): return len(self.particles)
    def getParticleParameters(self, i): return list(self.particles[i])
    def setParticleParameters(self, i, q, s, e): self.particles[i] = [q, s, e]
    def getNumExceptions(self): return len(self.exceptions)
    def getExceptionParameters(self, i): return list(self.exceptions[i])
    def setExceptionParameters(self, i, a, b, q, s, e): self.exceptions[i] = [a, b, q, s, e]
    def addException(self, a, b, q, s, e, replace=False): self.exceptions.append([a, b, q, s, e])
    def getNonbondedMethod(self): return self.method
    def setNonbondedMethod(self, m): self.method = m
    def getCutoffDistance(self): return self.cutoff
    def setCutoffDistance(self, c): self.cutoff = c
    def getUseSwitchingFunction(self): return self.use_switch
    def setUseSwitchingFunction(self, f): self.use_switch = f
    def getSwitchingDistance(self): return self.switch_distance
    def setSwitchingDistance(self, d): self.switch_distance = d
    def getUseDispersionCorrection(self): return self.dispersion
    def setUseDispersionCorrection(self, f): self.dispersion = f
    def getReactionFieldDielectric(self): return self.rf_dielectric
    def getEwaldErrorTolerance(self): return self.ewald_tol
    def addParticleParameterOffset(self, *a): self.offsets.append(a)
    def addExceptionParameterOffset(self, *a): self.offsets.append(a)


class CustomNonbondedForce(_Force):
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic = range(3)

    def __init__(self, expression):
        super().__init__()
        self.expression = expression
        self.per_particle, self.particles, self.groups, self.exclusions = [], [], [], []
        self.method = self.cutoff = self.use_switch = self.switch_distance = self.lrc = None

    def addPerParticleParameter(self, n): self.per_particle.append(n)
    def addParticle(self, p): self.particles.append(list(p))
    def addInteractionGroup(self, a, b): self.groups.append((sorted(a), sorted(b)))
    def addExclusion(self, a, b): self.exclusions.append((a, b))
    def setNonbondedMethod(self, m): self.method = m
    def setCutoffDistance(self, c): self.cutoff = c
    def setUseSwitchingFunction(self, f): self.use_switch = f
    def setSwitchingDistance(self, d): self.switch_distance = d
    def setUseLongRangeCorrection(self, f): self.lrc = f


class CustomBondForce(_Force):
    def __init__(self, expression):
        super().__init__()
        self.expression, self.per_bond, self.bonds = expression, [], []

    def addPerBondParameter(self, n): self.per_bond.append(n)
    def addBond(self, a, b, p): self.bonds.append((a, b, list(p)))


def load_factory():
    openmm = types.ModuleType('openmm')
    openmm.NonbondedForce, openmm.CustomNonbondedForce, openmm.CustomBondForce = NonbondedForce, CustomNonbondedForce, CustomBondForce
    src = open('/root/reference/openmmtools/alchemy/alchemy.py').read()
    tree = ast.parse(src)
    keep = []
    for n in tree.body:
        if isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == '_ALCHEMICAL_REGION_ARGS' for t in n.targets):
            keep.append(n)
        if isinstance(n, ast.ClassDef) and n.name in ('AlchemicalRegion', 'AbsoluteAlchemicalFactory'):
            keep.append(n)
        if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Attribute) and ast.unparse(n.targets[0]) == 'AlchemicalRegion.__new__.__defaults__':
            keep.append(n)
    ns = {'openmm': openmm, 'unit': u, 'np': np, 'copy': copy, 'collections': collections, 'itertools': itertools,
          're': re, 'logging': logging, 'logger': logging.getLogger('ref_alchemy'),
          'ONE_4PI_EPS0': 138.935456, 'utils': types.SimpleNamespace(), 'states': types.SimpleNamespace(),
          'forcefactories': types.SimpleNamespace()}
    exec(compile(ast.Module(body=keep, type_ignores=[]), 'ref_alchemy', 'exec'), ns)
    return ns['AbsoluteAlchemicalFactory'], ns['AlchemicalRegion']


# ------------------------------------------------------------------------------------------------- numpy evaluation
def min_image(d, L):
    return d - L * np.round(d / L)


def nonbonded_energy(force, x, L):
    """NonbondedForce LJ part (charges are zero here): Lorentz-Berthelot, periodic cutoff, no switch."""
    rc = float(u.to_md(force.getCutoffDistance()))
    sig = np.array([float(u.to_md(p[1])) for p in force.particles]); eps = np.array([float(u.to_md(p[2])) for p in force.particles])
    U = 0.0
    for i in range(len(sig)):
        for j in range(i + 1, len(sig)):
            if eps[i] == 0.0 or eps[j] == 0.0:
                continue
            d = min_image(x[i] - x[j], L); r = np.sqrt(d @ d)
            if r < rc:
                s = 0.5 * (sig[i] + sig[j]); e = np.sqrt(eps[i] * eps[j])
                U += 4.0 * e * ((s / r) ** 12 - (s / r) ** 6)
    return U


def custom_nonbonded_energy(force, x, L, overrides):
    rc = float(u.to_md(force.cutoff))
    g = dict(force.globals); g.update({k: v for k, v in overrides.items() if k in g})
    pairs = set()
    for a, b in force.groups:
        for i in a:
            for j in b:
                if i != j:
                    pairs.add((min(i, j), max(i, j)))
    U = 0.0
    for i, j in sorted(pairs):
        d = min_image(x[i] - x[j], L); r = np.sqrt(d @ d)
        if r >= rc:
            continue
        names = dict(g, r=r)
        for k, name in enumerate(force.per_particle):
            names[name + '1'] = float(u.to_md(force.particles[i][k])); names[name + '2'] = float(u.to_md(force.particles[j][k]))
        U += eval_lepton_program(force.expression, names)
    return U


def describe(force):
    d = {'type': type(force).__name__, 'globals': dict(force.globals)}
    if isinstance(force, CustomNonbondedForce):
        d.update(expression=force.expression, per_particle=force.per_particle, n_particles=len(force.particles),
                 groups=[[len(a), len(b), a[:3], b[:3]] for a, b in force.groups], method=force.method,
                 cutoff=float(u.to_md(force.cutoff)), use_switch=bool(force.use_switch),
                 switch_distance=float(u.to_md(force.switch_distance)), lrc=bool(force.lrc), n_exclusions=len(force.exclusions))
    elif isinstance(force, CustomBondForce):
        d.update(expression=force.expression, per_bond=force.per_bond, n_bonds=len(force.bonds))
    else:
        d.update(method=force.method, cutoff=float(u.to_md(force.cutoff)), use_switch=bool(force.use_switch),
                 dispersion=bool(force.dispersion),
                 eps=[float(u.to_md(p[2])) for p in force.particles], sigma=[float(u.to_md(p[1])) for p in force.particles],
                 charge=[float(u.to_md(p[0])) for p in force.particles])
    return d


N, N_ALCH = 40, 6
LAMBDAS = [0.0, 0.3, 0.7, 1.0]
CONFIGS = [  # annihilate_sterics, disable_alchemical_dispersion_correction, softcore (alpha, a, b, c)
    (False, False, (0.5, 1, 1, 6)),
    (True, False, (0.5, 1, 1, 6)),
    (False, True, (0.3, 2, 1.5, 12)),
]


def reference_force(s):
    """The NonbondedForce of testsystems.LennardJonesFluid (testsystems.py:1956-1989), switching function off."""
    f = NonbondedForce()
    f.setNonbondedMethod(NonbondedForce.CutoffPeriodic)
    f.setCutoffDistance(s['rc'] * u.nanometer)
    f.setUseSwitchingFunction(False)
    f.setSwitchingDistance(s['rs'] * u.nanometer)
    f.setUseDispersionCorrection(True)
    for i in range(s['N']):
        f.addParticle(0.0 * u.elementary_charge, s['sigma'][i] * u.nanometer, s['eps'][i] * u.kilojoules_per_mole)
    return f


if __name__ == '__main__':
    Factory, Region = load_factory()
    s = lj_setup(N=N, n_alch=N_ALCH, reduced_density=0.4, seed=77)
    x, L = s['x'], s['L']
    out = {'x': x, 'L': L}
    for c, (annihilate, disable_lrc, (alpha, a, b, cc)) in enumerate(CONFIGS):
        factory = Factory(disable_alchemical_dispersion_correction=disable_lrc)
        region = Region(alchemical_atoms=list(range(N_ALCH)), annihilate_sterics=annihilate, softcore_alpha=alpha,
                        softcore_a=a, softcore_b=b, softcore_c=cc)
        forces_by_lambda = factory._alchemically_modify_NonbondedForce(reference_force(s), [region], frozenset())
        desc = {k: [describe(f) for f in v] for k, v in forces_by_lambda.items()}
        out['config%d_forces' % c] = np.array(json.dumps(desc))
        U = []
        for lam in LAMBDAS:
            e = 0.0
            for key, forces in forces_by_lambda.items():
                for f in forces:
                    if isinstance(f, NonbondedForce):
                        e += nonbonded_energy(f, x, L)
                    elif isinstance(f, CustomNonbondedForce):
                        e += custom_nonbonded_energy(f, x, L, {'lambda_sterics': lam, 'lambda_electrostatics': lam})
                    else:
                        assert len(f.bonds) == 0
            U.append(e)
        out['config%d_U' % c] = np.array(U)
        print(c, annihilate, disable_lrc, U)
        for k, v in desc.items():
            for f in v:
                print('   ', repr(k), f['type'], {q: f[q] for q in ('groups', 'lrc', 'use_switch', 'globals') if q in f})
    dst = os.path.join(HERE, 'alchemy_golden.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, os.path.getsize(dst))
