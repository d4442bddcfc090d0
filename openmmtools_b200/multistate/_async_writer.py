"""Write-behind proxy for a MultiStateReporter: the records of iteration n go to storage on a writer thread while the
device already runs iteration n+1 (the engine calls release the GIL).  The reference writes synchronously
(multistatesampler.py:1191-1207); SURVEY.md 8 f-1 asks for the overlap.

Every `write_*` call is queued with a private copy of its arguments (the sampler's matrices are page-locked buffers that
the engine overwrites in the next iteration; SamplerStates are views of the host store) and executed in order, so the
commit marker (`write_last_iteration`) still reaches the storage after the records it commits.  Anything else -- reads,
`close`, `sync`, attribute access -- first drains the queue.  An exception raised by a queued write is re-raised by the
next call on the proxy."""
import copy
import queue
import threading

import numpy as np


def _private(x):
    if isinstance(x, np.ndarray):
        return np.array(x, copy=True)
    if isinstance(x, (list, tuple)):
        return type(x)(_private(v) for v in x)
    if isinstance(x, dict):
        return {k: _private(v) for k, v in x.items()}
    if isinstance(x, (int, float, str, bool, type(None))):
        return x
    return copy.deepcopy(x)


class AsyncReporter:
    def __init__(self, reporter, max_pending=64):
        object.__setattr__(self, '_reporter', reporter)
        object.__setattr__(self, '_queue', queue.Queue(maxsize=max_pending))
        object.__setattr__(self, '_error', None)
        t = threading.Thread(target=self._run, name='rx-reporter-writer', daemon=True)
        object.__setattr__(self, '_thread', t)
        t.start()

    def _run(self):
        while True:
            item = self._queue.get()
            try:
                if item is None:
                    return
                if self._error is None:
                    fn, args, kwargs = item
                    fn(*args, **kwargs)
            except BaseException as e:       # kept for the caller: a daemon thread must not die silently
                object.__setattr__(self, '_error', e)
            finally:
                self._queue.task_done()

    def drain(self):
        """Block until every queued write is on storage; re-raise a writer error."""
        self._queue.join()
        if self._error is not None:
            e = self._error
            object.__setattr__(self, '_error', None)
            raise e

    # configuration that never changes while a run is in progress: read without draining
    _PASS_THROUGH = ('checkpoint_interval', 'wants_analysis_states', 'filepath', 'is_open', 'storage_exists')

    def __getattr__(self, name):
        target = getattr(self._reporter, name)
        if name in self._PASS_THROUGH:
            return target
        if name.startswith('write_') and callable(target):
            def enqueue(*args, **kwargs):
                if self._error is not None:
                    self.drain()
                self._queue.put((target, _private(args), _private(kwargs)))
            return enqueue
        self.drain()
        return target

    def __setattr__(self, name, value):
        self.drain()
        setattr(self._reporter, name, value)

    def close(self):
        self.drain()
        return self._reporter.close()

    def shutdown(self):
        """Drain and stop the writer thread (the reporter itself stays open)."""
        self.drain()
        self._queue.put(None)
        self._thread.join()
