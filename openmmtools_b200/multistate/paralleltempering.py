"""ParallelTemperingSampler: replica exchange among temperatures (T-REMD) on the B200 engine.

Mirrors /root/reference/openmmtools/multistate/paralleltempering.py: ``create(thermodynamic_state, sampler_states,
storage, min_temperature, max_temperature, n_temperatures, temperatures)`` builds log-spaced
temperatures (``np.logspace``, :162), deep-copies the state per temperature (:167-170) and defers to the replica-exchange
``create``.  The reference's O(K) energy shortcut (``_compute_replica_energies``, :175-237: u[k,l] = beta_l U(x_k))
is what the engine's energy kernel does for every system anyway (one potential evaluation per replica, then one
multiply per state), so nothing is overridden here.
"""
import copy
from .replicaexchange import ReplicaExchangeSampler
from .. import unit


class ParallelTemperingSampler(ReplicaExchangeSampler):
    _TITLE_TEMPLATE = ('Parallel tempering simulation created using ParallelTemperingSampler '
                       'class of openmmtools_b200.multistate on {}')

    @staticmethod
    def _temperature_ladder(tmin, tmax, n):
        """paralleltempering.py:162: np.logspace(log10(Tmin), log10(Tmax), num=n), in kelvin."""
        import numpy as np
        return [float(t) for t in np.logspace(np.log10(tmin), np.log10(tmax), num=n)]

    def create(self, thermodynamic_state, sampler_states, storage=None, min_temperature=None, max_temperature=None,
               n_temperatures=None, temperatures=None, **kwargs):
        if not isinstance(sampler_states, (list, tuple)):
            sampler_states = [sampler_states]
        if temperatures is not None:
            if any(v is not None for v in (min_temperature, max_temperature, n_temperatures)):
                raise ValueError("Cannot set both 'temperatures' and 'min_temperature', 'max_temperature', "
                                 "and 'n_temperatures' at the same time.")
            temperatures = [float(unit.to_md(t, unit.kelvin, 'temperature')) for t in temperatures]
        elif all(v is not None for v in (min_temperature, max_temperature, n_temperatures)):
            tmin = float(unit.to_md(min_temperature, unit.kelvin, 'min_temperature'))
            tmax = float(unit.to_md(max_temperature, unit.kelvin, 'max_temperature'))
            temperatures = self._temperature_ladder(tmin, tmax, int(n_temperatures))
        else:
            raise ValueError("Either 'temperatures' or all of 'min_temperature', 'max_temperature', and "
                             "'n_temperatures' must be provided.")
        thermodynamic_states = [copy.deepcopy(thermodynamic_state) for _ in temperatures]
        for state, temperature in zip(thermodynamic_states, temperatures):
            state.temperature = temperature * unit.kelvin
        super().create(thermodynamic_states, sampler_states, storage=storage, **kwargs)
