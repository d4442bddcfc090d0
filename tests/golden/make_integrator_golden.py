#!/usr/bin/env python
"""Golden vectors for the Langevin splitting integrator and the soft-core sterics expression, from the REAL
reference code (build container only).

1. ``openmmtools/integrators.py`` is executed unmodified against a *recording* stand-in for ``openmm.CustomIntegrator``
   (OpenMM itself is not installable here): ``LangevinIntegrator(...)`` then yields its own step program -- the
   ``addComputePerDof`` expression strings, the global constants a, b, kT -- exactly as it would hand it to OpenMM.
   The program is interpreted with numpy in float64 under CustomIntegrator's documented semantics (``f`` is the force
   at the current positions, ``gaussian`` a fresh standard normal per degree of freedom, if-blocks on globals), with
   forces from the oracle's LJ model and the Gaussians injected, and the final positions/velocities are stored.
2. ``AbsoluteAlchemicalFactory._get_sterics_energy_expressions`` is lifted by AST and its Lepton expression strings are
   evaluated on a grid of (r, lambda, sigma_i, epsilon_i, softcore parameters).

Output: tests/golden/integrator_golden.npz
"""
import ast, importlib.util, os, re, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, HERE)
from openmmtools_b200 import unit as u, constants
from helpers import lj_setup, oracle_system, KB


# ---------------------------------------------------------------------------------------------- recording stand-in
class CustomIntegrator:
    def __init__(self, timestep):
        self.dt = float(u.to_md(timestep))
        self.globals, self.gorder, self.perdof, self.program = {}, [], {}, []

    def addGlobalVariable(self, name, value):
        self.globals[name] = float(u.to_md(value)); self.gorder.append(name); return len(self.gorder) - 1

    def addPerDofVariable(self, name, value): self.perdof[name] = float(value)
    def addComputePerDof(self, var, expr): self.program.append(('perdof', var, expr))
    def addComputeGlobal(self, var, expr): self.program.append(('global', var, expr))
    def addComputeSum(self, var, expr): self.program.append(('sum', var, expr))
    def addConstrainPositions(self): self.program.append(('constrain_x',))
    def addConstrainVelocities(self): self.program.append(('constrain_v',))
    def addUpdateContextState(self): self.program.append(('update_context',))
    def beginIfBlock(self, cond): self.program.append(('if', cond))
    def endBlock(self): self.program.append(('end',))
    def getNumGlobalVariables(self): return len(self.gorder)
    def getGlobalVariableName(self, i): return self.gorder[i]
    def getGlobalVariableByName(self, n): return self.globals[n]
    def setGlobalVariableByName(self, n, v): self.globals[n] = float(u.to_md(v))
    def setConstraintTolerance(self, t): self.tol = t
    def getStepSize(self): return self.dt


def load_reference_integrators():
    mm = types.ModuleType('openmm'); mm.CustomIntegrator = CustomIntegrator; mm.unit = u
    sys.modules['openmm'] = mm; sys.modules['openmm.unit'] = u
    pkg = types.ModuleType('openmmtools'); pkg.__path__ = []; sys.modules['openmmtools'] = pkg
    cst = types.ModuleType('openmmtools.constants'); cst.kB = constants.kB
    sys.modules['openmmtools.constants'] = cst; pkg.constants = cst

    class Restorable:
        def __init__(self, *a, **k): super().__init__(*a, **k)
    ut = types.ModuleType('openmmtools.utils'); ut.RestorableOpenMMObject = Restorable
    sys.modules['openmmtools.utils'] = ut; pkg.utils = ut
    rs = types.ModuleType('openmmtools.respa'); rs.MTSIntegrator = type('MTSIntegrator', (), {})
    sys.modules['openmmtools.respa'] = rs; pkg.respa = rs
    spec = importlib.util.spec_from_file_location('ref_integrators', '/root/reference/openmmtools/integrators.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def lepton(expr, names):
    """Evaluate one Lepton arithmetic expression (the subset the reference uses: + - * / ^ sqrt and names)."""
    return eval(expr.replace('^', '**'), {'__builtins__': {}, 'sqrt': np.sqrt}, names)


def interpret(integ, x, v, mass, force_fn, noise, n_steps):
    """Run the recorded CustomIntegrator program n_steps times (no constraints in these systems)."""
    g = dict(integ.globals)
    per = {k: np.full_like(x, val) for k, val in integ.perdof.items()}
    m = mass[:, None]
    o = 0
    f_cache = [None]

    def names():
        d = dict(g); d.update(per)
        d.update(x=x, v=v, m=m, dt=integ.dt)
        return d

    class Lazy(dict):   # 'f' and 'gaussian' are materialised when an expression mentions them
        def __missing__(self, key):
            if key == 'f':
                if f_cache[0] is None:
                    f_cache[0] = force_fn(x)
                return f_cache[0]
            raise KeyError(key)

    for _ in range(n_steps):
        skip = 0
        for ins in integ.program:
            if ins[0] == 'end':
                skip = max(0, skip - 1); continue
            if ins[0] == 'if':
                lhs, rhs = [s.strip() for s in ins[1].split('=')]
                if skip or not (g[lhs] == float(rhs)):
                    skip += 1
                continue
            if skip:
                continue
            if ins[0] == 'perdof':
                var, expr = ins[1], ins[2]
                d = Lazy(names())
                if re.search(r'\bgaussian\b', expr):
                    d['gaussian'] = noise[o]; o += 1
                val = np.array(lepton(expr, d), dtype=np.float64) + np.zeros_like(x)
                if var == 'x':
                    x = val; f_cache[0] = None
                elif var == 'v':
                    v = val
                else:
                    per[var] = val
            elif ins[0] == 'global':
                g[ins[1]] = float(lepton(ins[2], names()))
            elif ins[0] in ('constrain_x', 'constrain_v', 'update_context'):
                pass
            else:
                raise NotImplementedError(ins)
    assert o == len(noise)
    return x, v


CASES = [  # splitting, n_steps, dt (ps), gamma (1/ps), T (K), lambda
    ('V R O R V', 5, 0.002, 10.0, 300.0, 1.0),
    ('V R O R V', 4, 0.001, 1.0, 350.0, 0.4),
    ('O V R V O', 4, 0.002, 10.0, 300.0, 0.7),
    ('R V O', 3, 0.002, 5.0, 320.0, 0.0),
    ('V R R O R R V', 3, 0.002, 10.0, 300.0, 0.25),
    ('O R V R O', 4, 0.0015, 20.0, 280.0, 0.9),
]


def case_inputs(idx, N=48):
    s = lj_setup(N=N, n_alch=5, seed=100 + idx)
    rng = np.random.default_rng(500 + idx)
    v0 = rng.normal(scale=0.3, size=(N, 3))
    splitting, n_steps = CASES[idx][0], CASES[idx][1]
    noise = rng.normal(size=(n_steps * splitting.split().count('O'), N, 3))
    return s, v0, noise


def sterics_expressions():
    tree = ast.parse(open('/root/reference/openmmtools/alchemy/alchemy.py').read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'AbsoluteAlchemicalFactory'][0]
    fn = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == '_get_sterics_energy_expressions'][0]
    fn.decorator_list = []
    ns = {}
    exec(ast.unparse(fn), ns)
    return ns['_get_sterics_energy_expressions'](None, [''])


def eval_lepton_program(text, names):
    """'result; a = ...; b = ...;' -- definitions may come after their use."""
    parts = [p.strip() for p in text.split(';') if p.strip()]
    defs = {}
    result = None
    for p in parts:
        if re.match(r'^[A-Za-z_][A-Za-z_0-9]*\s*=', p):
            k, e = p.split('=', 1); defs[k.strip()] = e.strip()
        else:
            result = p

    class Resolver(dict):
        def __missing__(self, key):
            if key in defs:
                val = lepton(defs[key], self); self[key] = val; return val
            raise KeyError(key)
    return lepton(result, Resolver(names))


STERICS_GRID = dict(
    r=np.array([0.18, 0.25, 0.31, 0.34, 0.3816, 0.45, 0.6, 0.9]),
    lam=np.array([0.0, 0.1, 0.5, 0.9, 1.0]),
    params=[  # sigma1, sigma2, eps1, eps2, alpha, a, b, c
        (0.34, 0.34, 0.995792, 0.995792, 0.5, 1.0, 1.0, 6.0),
        (0.30, 0.40, 0.5, 1.2, 0.5, 1.0, 1.0, 6.0),
        (0.34, 0.25, 0.8, 0.3, 0.3, 2.0, 1.5, 12.0),
        (0.36, 0.36, 1.0, 1.0, 0.7, 1.0, 2.0, 4.0),
    ])


if __name__ == '__main__':
    out = {}
    mod = load_reference_integrators()
    for idx, (splitting, n_steps, dt, gamma, T, lam) in enumerate(CASES):
        s, v0, noise = case_inputs(idx)
        integ = mod.LangevinIntegrator(temperature=T * u.kelvin, collision_rate=gamma / u.picoseconds,
                                       timestep=dt * u.picoseconds, splitting=splitting)
        osys = oracle_system(s)
        x, v = interpret(integ, s['x'].copy(), v0.copy(), s['mass'], lambda xx: osys.energy(xx, lam, forces=True)[2],
                         noise, n_steps)
        out['case%d_x' % idx] = x; out['case%d_v' % idx] = v
        out['case%d_globals' % idx] = np.array([integ.globals['kT'], integ.globals['a'], integ.globals['b']])
        if idx == 0:
            out['program_VRORV'] = np.array(['|'.join(map(str, p)) for p in integ.program])
    mixing, energy = sterics_expressions()
    out['sterics_expression'] = np.array([mixing, energy])
    G = STERICS_GRID
    U = np.zeros((len(G['params']), len(G['lam']), len(G['r'])))
    for p, (s1, s2, e1, e2, alpha, a, b, c) in enumerate(G['params']):
        for l, lam in enumerate(G['lam']):
            for q, r in enumerate(G['r']):
                U[p, l, q] = eval_lepton_program(energy + mixing, dict(
                    r=r, lambda_sterics=lam, sigma1=s1, sigma2=s2, epsilon1=e1, epsilon2=e2, softcore_alpha=alpha,
                    softcore_a=a, softcore_b=b, softcore_c=c))
    out['sterics_U'] = U
    dst = os.path.join(HERE, 'integrator_golden.npz')
    np.savez_compressed(dst, **out)
    print('wrote', dst, os.path.getsize(dst))
    print(out['program_VRORV'])
    print(out['sterics_expression'])
