"""multistate/_async_writer.AsyncReporter: write-behind of the per-iteration records -- order, private
copies, draining on reads, error propagation; on the GPU a run with asynchronous reporting leaves exactly the storage of a
synchronous one."""
import os
import threading
import time
import numpy as np
import pytest
from openmmtools_b200.multistate._async_writer import AsyncReporter


class Recorder:
    checkpoint_interval = 5

    def __init__(self, delay=0.0, fail_at=None):
        self.calls, self.delay, self.fail_at = [], delay, fail_at
        self.writer_threads = set()

    def write_energies(self, u, it):
        time.sleep(self.delay)
        self.writer_threads.add(threading.get_ident())
        if it == self.fail_at:
            raise IOError('disk full')
        self.calls.append(('energies', it, float(u.sum())))

    def write_last_iteration(self, it):
        self.calls.append(('last', it))

    def read_last_iteration(self):
        return max([c[1] for c in self.calls if c[0] == 'last'], default=None)

    def wants_analysis_states(self, it):
        return False


def test_writes_are_ordered_private_and_behind():
    rec = Recorder(delay=0.02)
    a = AsyncReporter(rec)
    u = np.ones((4, 4))
    t0 = time.time()
    for it in range(1, 6):
        u[:] = it                         # the caller's buffer is overwritten by the next iteration
        a.write_energies(u, it)
        a.write_last_iteration(it)
    assert time.time() - t0 < 0.05        # the caller did not wait for the 5 x 20 ms of writes
    assert a.checkpoint_interval == 5     # configuration reads do not drain
    assert len(rec.calls) < 10
    assert a.read_last_iteration() == 5   # a read drains first
    assert rec.calls == [c for it in range(1, 6) for c in (('energies', it, 16.0 * it), ('last', it))]
    assert rec.writer_threads == {a._thread.ident} and a._thread.ident != threading.get_ident()
    a.shutdown()
    assert not a._thread.is_alive()


def test_writer_error_reaches_the_caller_and_nothing_is_committed_after_it():
    rec = Recorder(fail_at=3)
    a = AsyncReporter(rec)
    u = np.zeros((2, 2))
    for it in range(1, 6):
        try:
            a.write_energies(u, it)
            a.write_last_iteration(it)
        except IOError:
            break
    with pytest.raises(IOError):
        a.drain()
        a.drain()                          # (whichever call sees it first)
        raise IOError('already delivered')
    committed = [c[1] for c in rec.calls if c[0] == 'last']
    assert committed == [1, 2]             # iteration 3 and everything queued behind the failure was dropped
    a.shutdown()


@pytest.mark.gpu
def test_asynchronous_reporting_leaves_the_same_storage(tmp_path):
    from openmmtools_b200 import testsystems, states, mcmc, multistate, unit
    def run(path, asynchronous):
        ho = testsystems.HarmonicOscillator()
        ts = [states.ThermodynamicState(ho.system, T * unit.kelvin) for T in (300, 320, 340, 360)]
        s = multistate.ReplicaExchangeSampler(mcmc_moves=mcmc.LangevinSplittingDynamicsMove(n_steps=20), number_of_iterations=12, seed=5)
        s.asynchronous_reporting = asynchronous
        rep = multistate.MultiStateReporter(str(path), checkpoint_interval=4)
        s.create(ts, [states.SamplerState(ho.positions)], storage=rep)
        s.run()
        assert s.asynchronous_reporting == asynchronous
        r = multistate.MultiStateReporter(str(path)); r.open('r')
        out = dict(last=r.read_last_iteration(last_checkpoint=False), energies=[r.read_energies(i)[0] for i in range(13)],
                   states=[r.read_replica_thermodynamic_states(i) for i in range(13)],
                   mix=[r.read_mixing_statistics(i) for i in range(1, 13)],
                   x=[st.positions for st in r.read_sampler_states(12)])
        r.close()
        return out
    a, b = run(tmp_path / 'sync', False), run(tmp_path / 'async', True)
    assert a['last'] == b['last'] == 12
    for k in ('energies', 'states'):
        for p, q in zip(a[k], b[k]):
            assert np.array_equal(np.asarray(p), np.asarray(q))
    for (na, npr), (ma, mpr) in zip(a['mix'], b['mix']):
        assert np.array_equal(na, ma) and np.array_equal(npr, mpr)
    for p, q in zip(a['x'], b['x']):
        assert np.array_equal(np.asarray(p.value_in_unit(unit.nanometer)), np.asarray(q.value_in_unit(unit.nanometer)))
