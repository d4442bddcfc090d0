"""ReplicaExchangeSampler: Gibbs / Metropolis swaps between replicas, on the GPU.

Mirrors /root/reference/openmmtools/multistate/replicaexchange.py (:52-447): ``replica_mixing_scheme`` in
{'swap-all', 'swap-neighbors', None} (:220-234), ``create`` round-robins sampler states and rejects more sampler
states than thermodynamic states (:239-253), ``_mix_replicas`` (:255-292) zeroes the count matrices and attempts
K^3 swaps ('swap-all') or one pass over neighbouring state pairs ('swap-neighbors').  The static
``_mix_all_replicas_numba(nswap, K, states, u, n_accepted, n_proposed)`` (:294-349) keeps its signature and
in-place semantics but runs the CUDA kernel.
"""
import numpy as np
from .multistatesampler import MultiStateSampler
from .. import _lib
from .._engine import Engine


class ReplicaExchangeSampler(MultiStateSampler):
    _TITLE_TEMPLATE = ('Replica-exchange sampler simulation created using ReplicaExchangeSampler class '
                       'of openmmtools_b200.multistate on {}')

    def __init__(self, replica_mixing_scheme='swap-all', **kwargs):
        super().__init__(**kwargs)
        self.replica_mixing_scheme = replica_mixing_scheme

    @property
    def replica_mixing_scheme(self):
        return self._replica_mixing_scheme

    @replica_mixing_scheme.setter
    def replica_mixing_scheme(self, value):
        supported = ['swap-all', 'swap-neighbors', None]     # replicaexchange.py:223-227
        if value not in supported:
            raise ValueError("Unknown replica mixing scheme '{}'. Supported values are {}.".format(value, supported))
        if self.locality is not None and value != 'swap-neighbors':
            raise ValueError("replica_mixing_scheme must be 'swap-neighbors' if locality is used")
        self._replica_mixing_scheme = value

    def _pre_write_create(self, thermodynamic_states, sampler_states, *args, **kwargs):
        n_states = len(thermodynamic_states)
        if len(sampler_states) > n_states:
            raise ValueError('Passed {} SamplerStates but only {} ThermodynamicStates'.format(
                len(sampler_states), n_states))
        sampler_states = [sampler_states[i % len(sampler_states)] for i in range(n_states)]
        super()._pre_write_create(thermodynamic_states, sampler_states, *args, **kwargs)

    def _mix_replicas(self):
        """Attempt to swap replicas according to the scheme; uses the energies of the previous iteration."""
        e = self._engine
        # the engine writes straight into the sampler's (page-locked) matrices
        out = (self._replica_thermodynamic_states, self._n_accepted_matrix, self._n_proposed_matrix)
        if self.replica_mixing_scheme == 'swap-neighbors':
            e.mix_swap_neighbors(out=out)
        elif self.replica_mixing_scheme == 'swap-all':
            e.mix_swap_all(self.n_replicas ** 3, out=out)
        else:
            assert self.replica_mixing_scheme is None
            self._n_accepted_matrix[:, :] = 0
            self._n_proposed_matrix[:, :] = 0
        return self._replica_thermodynamic_states

    # reference-compatible static kernel entry point (called directly by the reference's tests/test_mixing.py:41-43)
    _static_seed = [None]

    @staticmethod
    def _mix_all_replicas_numba(nswap_attempts, n_replicas, _replica_thermodynamic_states,
                                _energy_thermodynamic_states, _n_accepted_matrix, _n_proposed_matrix, seed=None):
        """In-place swap-all mixing on the GPU with the reference's argument order.  ``seed`` (extension) seeds the
        MT19937 stream like a jitted ``np.random.seed(seed)``; without it a random seed is used once per process and
        the stream continues across calls, as numba's thread-local generator does."""
        K = int(n_replicas)
        cache = ReplicaExchangeSampler.__dict__['_static_engines']
        eng = cache.get(K)
        if eng is None:
            eng = cache[K] = Engine(_lib.RX_SYSTEM_NONE, K, K)
            eng.mix_seed(int(np.random.SeedSequence().entropy & 0xFFFFFFFF) if seed is None else seed)
        elif seed is not None:
            eng.mix_seed(seed)
        eng.set_energies(np.ascontiguousarray(_energy_thermodynamic_states, dtype=np.float64))
        eng.set_replica_states(np.asarray(_replica_thermodynamic_states, dtype=np.int64))
        st, nacc, nprop = eng.mix_swap_all(int(nswap_attempts))
        _replica_thermodynamic_states[:] = st
        _n_accepted_matrix += nacc.astype(_n_accepted_matrix.dtype)
        _n_proposed_matrix += nprop.astype(_n_proposed_matrix.dtype)

    _static_engines = {}
