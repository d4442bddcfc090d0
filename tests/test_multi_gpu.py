"""Multi-process paths.  CPU (gloo, world_size 2): the host-side plumbing (rank/shard arithmetic, id broadcast).
GPU (marked gpu, needs >= 2 devices): two ranks shard the replicas, all-gather energy rows over NCCL and must
reproduce the single-GPU trajectory exactly (noise is keyed by global replica id, mixing is replicated)."""
import os
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
dist.init_process_group('gloo')
from openmmtools_b200._dist import TorchCommunicator
c = TorchCommunicator()
payload = bytes(range(128)) if c.rank == 0 else None
got = c.bcast_bytes(payload, 128)
assert got == bytes(range(128)), got
K = 10
k0 = (c.rank * K) // c.world_size; k1 = ((c.rank + 1) * K) // c.world_size
import torch
t = torch.tensor([float(k1 - k0)]); dist.all_reduce(t)
assert t.item() == K
g = c.gather_object({'rank': c.rank, 'rows': list(range(k0, k1))})
if c.rank == 0:
    assert [d['rank'] for d in g] == [0, 1] and sum(len(d['rows']) for d in g) == K
else:
    assert g is None
c.barrier()
print('ok', c.rank, k0, k1)
'''


def test_gloo_world_size_2_host_plumbing(tmp_path):
    script = tmp_path / 'w.py'
    script.write_text(WORKER % {'root': ROOT})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29533', str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count('ok') == 2


GPU_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import numpy as np
import torch.distributed as dist
dist.init_process_group('gloo')
from test_gpu_sampler import lj_sampler
from openmmtools_b200._dist import TorchCommunicator
s, asys, lambdas = lj_sampler(K=16, seed=321, communicator=TorchCommunicator())
s.run(3)
u = s._energy_thermodynamic_states
perm = s._replica_thermodynamic_states
x = s._engine.get_positions()
np.savez(%(out)r + '_%%d.npz' %% dist.get_rank(), u=u, perm=perm, x=x, k0=s._engine.k0)
dist.barrier()
'''


@pytest.mark.gpu
def test_two_gpus_reproduce_single_gpu(tmp_path):
    import ctypes
    try:
        cuda = ctypes.CDLL('libcudart.so.12')
    except OSError:
        cuda = ctypes.CDLL('libcudart.so')
    n = ctypes.c_int()
    cuda.cudaGetDeviceCount(ctypes.byref(n))
    if n.value < 2:
        pytest.skip('needs 2 GPUs')
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_gpu_sampler import lj_sampler
    s, _, _ = lj_sampler(K=16, seed=321)
    s.run(3)
    u1 = s._energy_thermodynamic_states.copy(); p1 = s._replica_thermodynamic_states.copy(); x1 = s._engine.get_positions()
    script = tmp_path / 'g.py'
    out = str(tmp_path / 'res')
    script.write_text(GPU_WORKER % {'root': ROOT, 'out': out})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29534', str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(2):
        d = np.load(out + '_%d.npz' % rank)
        assert np.array_equal(d['perm'], p1)
        assert np.array_equal(d['u'], u1)
        k0 = int(d['k0'])
        assert np.array_equal(d['x'], x1[k0:k0 + d['x'].shape[0]])
