#!/usr/bin/env python
"""Golden data for the two test systems of the hot path, from the REAL reference code (build container only).

``TestSystem``, ``LennardJonesFluid``, ``HarmonicOscillator`` and ``subrandom_particle_positions`` are lifted by AST from
/root/reference/openmmtools/testsystems.py and executed on recording stand-ins for ``openmm.System`` /
``NonbondedForce`` / ``CustomExternalForce`` / ``app.Topology`` (OpenMM itself is not installable here), with the
reference's own ``sobol.py`` for the sub-random positions.  Recorded: particle parameters, masses, cutoff, switching
distance, dispersion-correction flag, box vectors, positions; the oscillator's energy expression and global parameters.

Output: tests/golden/testsystems_golden.npz
"""
import ast, importlib.util, json, os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
from openmmtools_b200 import unit as u
from make_alchemy_golden import NonbondedForce, CustomNonbondedForce, _Force


class System:
    def __init__(self):
        self.masses, self.forces, self.box = [], [], None

    def addParticle(self, m): self.masses.append(m); return len(self.masses) - 1
    def getNumParticles(self): return len(self.masses)
    def addForce(self, f): self.forces.append(f); return len(self.forces) - 1
    def setDefaultPeriodicBoxVectors(self, a, b, c): self.box = (a, b, c)
    def getDefaultPeriodicBoxVectors(self): return self.box


class CustomExternalForce(_Force):
    def __init__(self, expression):
        super().__init__()
        self.expression, self.particles = expression, []

    def addParticle(self, i, p): self.particles.append((i, list(p)))


class _Topology:
    def addChain(self): return object()
    def addResidue(self, *a): return object()
    def addAtom(self, *a): return object()


def load():
    openmm = types.ModuleType('openmm')
    openmm.System, openmm.NonbondedForce, openmm.CustomNonbondedForce, openmm.CustomExternalForce = \
        System, NonbondedForce, CustomNonbondedForce, CustomExternalForce
    app = types.SimpleNamespace(Topology=_Topology, Element=types.SimpleNamespace(getBySymbol=lambda s: s))
    spec = importlib.util.spec_from_file_location('ref_sobol', '/root/reference/openmmtools/sobol.py')
    sobol = importlib.util.module_from_spec(spec); spec.loader.exec_module(sobol)
    pkg = types.ModuleType('openmmtools'); pkg.__path__ = []; pkg.sobol = sobol
    sys.modules['openmmtools'] = pkg; sys.modules['openmmtools.sobol'] = sobol
    tree = ast.parse(open('/root/reference/openmmtools/testsystems.py').read())
    keep = [n for n in tree.body if (isinstance(n, ast.ClassDef) and n.name in ('TestSystem', 'LennardJonesFluid', 'HarmonicOscillator'))
            or (isinstance(n, ast.FunctionDef) and n.name in ('subrandom_particle_positions', 'halton_sequence'))]
    ns = {'openmm': openmm, 'unit': u, 'np': np, 'app': app, 'DEFAULT_EWALD_ERROR_TOLERANCE': 1e-5}
    exec(compile(ast.Module(body=keep, type_ignores=[]), 'ref_testsystems', 'exec'), ns)
    return ns


def md(q): return np.asarray(u.to_md(q), dtype=np.float64)


if __name__ == '__main__':
    ns = load()
    out = {}
    for tag, kw in (('lj512', dict(nparticles=512)), ('lj100_dense', dict(nparticles=100, reduced_density=0.3, switch_width=2.0 * u.angstroms, cutoff=9.0 * u.angstroms)),
                    ('lj64_noswitch', dict(nparticles=64, switch_width=None, dispersion_correction=False))):
        fluid = ns['LennardJonesFluid'](**kw)
        sysm = fluid.system
        nb = [f for f in sysm.forces if isinstance(f, NonbondedForce)][0]
        assert len(sysm.forces) == 1
        out[tag + '_mass'] = np.array([float(md(m)) for m in sysm.masses])
        out[tag + '_charge'] = np.array([float(md(p[0])) for p in nb.particles])
        out[tag + '_sigma'] = np.array([float(md(p[1])) for p in nb.particles])
        out[tag + '_epsilon'] = np.array([float(md(p[2])) for p in nb.particles])
        out[tag + '_box'] = np.array([[float(md(c)) for c in v] for v in sysm.box])
        out[tag + '_nb'] = np.array([nb.method, float(md(nb.cutoff)), float(nb.use_switch), float(md(nb.switch_distance)), float(nb.dispersion)])
        out[tag + '_positions'] = md(fluid.positions)
        print(tag, out[tag + '_box'][0, 0], out[tag + '_nb'], out[tag + '_sigma'][0], out[tag + '_epsilon'][0], out[tag + '_mass'][0], out[tag + '_positions'][:2])
    ho = ns['HarmonicOscillator']()
    f = ho.system.forces[0]
    out['ho_expression'] = np.array(f.expression)
    out['ho_globals'] = np.array(json.dumps(f.globals))
    out['ho_mass'] = np.array([float(md(m)) for m in ho.system.masses])
    out['ho_positions'] = md(ho.positions)
    print(f.expression, f.globals, out['ho_mass'])
    dst = os.path.join(HERE, 'testsystems_golden.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, os.path.getsize(dst))
