"""MCMC moves of the replica-exchange path: Langevin splitting dynamics on the CUDA engine.

Mirrors /root/reference/openmmtools/mcmc.py: ``LangevinSplittingDynamicsMove`` (:1175-1316),
``LangevinDynamicsMove`` (:1023-1172), ``BaseIntegratorMove.apply`` (:668-776) and ``IntegratorMoveError``
(:538-600).  The splitting grammar and its validation follow ``integrators.LangevinIntegrator``
(/root/reference/openmmtools/integrators.py:1319-1363,1474-1537): tokens R, V, O; force groups (``V0``),
Metropolization (``{ }``), shadow-work and heat bookkeeping need per-substep energy sums and are not provided.
GHMC/HMC/MC displacement/barostat moves (:1323-1917) are other kernels and out of scope.
"""
import numpy as np
from . import unit
from . import _backend
from ._engine import EngineError
from . import _lib


class IntegratorMoveError(Exception):
    """The integrator produced NaN (mcmc.py:538-600)."""

    def __init__(self, message, move, context=None):
        super().__init__(message)
        self.move = move
        self.context = context

    def serialize_error(self, path_files_prefix):
        """Dump what is available for post-mortem (the reference writes system/integrator/state XML)."""
        import json
        with open(path_files_prefix + '-move.json', 'w') as f:
            json.dump({k: str(v) for k, v in self.move.__getstate__().items()}, f)


def parse_splitting(splitting):
    """Validate a splitting string, returning it without spaces (integrators.py:1319-1402; behaviour per string pinned
    by tests/golden/splitting_golden.json).  ``V0`` is the plain ``V`` of a system whose forces all sit in group 0."""
    tokens = splitting.split()
    steps = []
    groups = set()
    for step in tokens:
        if step in ('{', '}'):
            raise NotImplementedError('Metropolized splittings ({ }) are not provided on the B200 path')
        if step[0] == 'V' and len(step) > 1:
            try:
                group = int(step[1:])
            except ValueError:
                raise ValueError('You must use an integer force group')
            if group > 31:
                raise ValueError('OpenMM only allows up to 32 force groups')
            groups.add(group)
            steps.append('V')
            continue
        if step not in ('R', 'V', 'O'):
            raise ValueError("Invalid step name '{}' used; valid step names are R, V, O".format(step))
        steps.append(step)
    if groups - {0}:
        raise NotImplementedError('multiple-time-step splittings (V<group>) are not provided on the B200 path')
    joined = ''.join(steps)
    for need in 'RVO':
        assert need in joined, 'splitting must contain R, V and O steps'
    return joined


class MCMCMove:
    """Interface: ``apply(thermodynamic_state, sampler_state, context_cache=None)`` mutates the sampler state
    (mcmc.py:143-175)."""

    def apply(self, thermodynamic_state, sampler_state, context_cache=None):
        raise NotImplementedError


class BaseIntegratorMove(MCMCMove):
    def __init__(self, n_steps, reassign_velocities=False, n_restart_attempts=4, context_cache=None):
        self.n_steps = n_steps
        self.reassign_velocities = reassign_velocities
        self.n_restart_attempts = n_restart_attempts
        self.context_cache = context_cache
        self._seed_counter = 0

    # engine-facing description: (timestep ps, collision rate 1/ps, n_steps, splitting)
    def _integrator_parameters(self):
        raise NotImplementedError

    def apply(self, thermodynamic_state, sampler_state, context_cache=None):
        """Propagate one replica (mcmc.py:668-776): builds a single-replica engine, steps, updates the state.
        NaN -> retry up to n_restart_attempts times, then IntegratorMoveError (mcmc.py:706-759)."""
        dt, gamma, n_steps, splitting = self._integrator_parameters()
        device = _backend.default_device(context_cache or self.context_cache)
        eng = _backend.build_engine([thermodynamic_state], 1, device=device)
        try:
            eng.set_integrator(dt, gamma, n_steps, splitting)
            has_v = sampler_state._velocities is not None
            reassign = self.reassign_velocities or not has_v
            attempts = 0
            seed0 = getattr(self, 'seed', None)
            if seed0 is None:
                seed0 = int(np.random.SeedSequence().entropy & 0xFFFFFFFFFFFFFFFF)
            while True:
                eng.set_positions(sampler_state._positions[None])
                if has_v:
                    eng.set_velocities(sampler_state._velocities[None])
                try:
                    eng.propagate(seed0 + attempts, self._seed_counter, reassign_velocities=reassign)
                    break
                except EngineError as e:
                    if e.code != _lib.RX_ERR_NAN:
                        raise
                    attempts += 1
                    if attempts > self.n_restart_attempts:
                        raise IntegratorMoveError('Potential energy is NaN after {} attempts of integration '
                                                  'with move {}'.format(attempts, self.__class__.__name__), self)
            self._seed_counter += 1
            pot, kin = eng.get_replica_energies()
            sampler_state._update(eng.get_positions()[0], eng.get_velocities()[0], pot[0], kin[0])
        finally:
            eng.close()

    def __getstate__(self):
        return dict(n_steps=self.n_steps, reassign_velocities=self.reassign_velocities,
                    n_restart_attempts=self.n_restart_attempts)

    def __setstate__(self, s):
        self.n_steps = s['n_steps']
        self.reassign_velocities = s['reassign_velocities']
        self.n_restart_attempts = s['n_restart_attempts']
        self.context_cache = None
        self._seed_counter = 0


class LangevinSplittingDynamicsMove(BaseIntegratorMove):
    """Langevin dynamics with an arbitrary R/V/O splitting (mcmc.py:1175-1316); default "V R O R V" (BAOAB)."""

    def __init__(self, timestep=1.0 * unit.femtosecond, collision_rate=10.0 / unit.picoseconds, n_steps=1000,
                 reassign_velocities=False, splitting="V R O R V", constraint_tolerance=1.0e-8,
                 measure_shadow_work=False, measure_heat=False, **kwargs):
        super().__init__(n_steps=n_steps, reassign_velocities=reassign_velocities, **kwargs)
        if measure_shadow_work or measure_heat:
            raise NotImplementedError('shadow-work / heat bookkeeping is not provided on the B200 path')
        parse_splitting(splitting)
        self.timestep = timestep
        self.collision_rate = collision_rate
        self.splitting = splitting
        self.constraint_tolerance = constraint_tolerance
        self.measure_shadow_work = measure_shadow_work
        self.measure_heat = measure_heat

    def _integrator_parameters(self):
        return (float(unit.to_md(self.timestep, unit.picosecond, 'timestep')),
                float(unit.to_md(self.collision_rate, unit.picosecond ** -1, 'collision_rate')),
                int(self.n_steps), parse_splitting(self.splitting))

    def __getstate__(self):
        s = super().__getstate__()
        s.update(timestep=self.timestep, collision_rate=self.collision_rate, splitting=self.splitting,
                 constraint_tolerance=self.constraint_tolerance, measure_shadow_work=self.measure_shadow_work,
                 measure_heat=self.measure_heat)
        return s

    def __setstate__(self, s):
        super().__setstate__(s)
        for k in ('timestep', 'collision_rate', 'splitting', 'constraint_tolerance', 'measure_shadow_work', 'measure_heat'):
            setattr(self, k, s[k])


class LangevinDynamicsMove(BaseIntegratorMove):
    """Langevin dynamics with OpenMM's LangevinMiddleIntegrator discretisation (mcmc.py:1023-1172): per step
    v += dt f/m ; x += dt/2 v ; v = a v + b sqrt(kT/m) xi ; x += dt/2 v, i.e. the splitting "V R O R"."""

    def __init__(self, timestep=1.0 * unit.femtosecond, collision_rate=10.0 / unit.picoseconds, n_steps=1000,
                 reassign_velocities=False, constraint_tolerance=1e-8, **kwargs):
        super().__init__(n_steps=n_steps, reassign_velocities=reassign_velocities, **kwargs)
        self.timestep = timestep
        self.collision_rate = collision_rate
        self.constraint_tolerance = constraint_tolerance

    def _integrator_parameters(self):
        return (float(unit.to_md(self.timestep, unit.picosecond, 'timestep')),
                float(unit.to_md(self.collision_rate, unit.picosecond ** -1, 'collision_rate')),
                int(self.n_steps), 'VROR')

    def __getstate__(self):
        s = super().__getstate__()
        s.update(timestep=self.timestep, collision_rate=self.collision_rate, constraint_tolerance=self.constraint_tolerance)
        return s

    def __setstate__(self, s):
        super().__setstate__(s)
        for k in ('timestep', 'collision_rate', 'constraint_tolerance'):
            setattr(self, k, s[k])


def same_integrator(a, b):
    return type(a) is type(b) and a._integrator_parameters() == b._integrator_parameters() and \
        a.reassign_velocities == b.reassign_velocities
