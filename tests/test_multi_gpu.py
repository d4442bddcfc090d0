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


def _two_devices():
    import ctypes
    try:
        cuda = ctypes.CDLL('libcudart.so.12')
    except OSError:
        cuda = ctypes.CDLL('libcudart.so')
    n = ctypes.c_int()
    cuda.cudaGetDeviceCount(ctypes.byref(n))
    return n.value >= 2


WIDE_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import numpy as np
import torch.distributed as dist
dist.init_process_group('gloo')
from openmmtools_b200._dist import TorchCommunicator
from test_multi_gpu import sams_run, tremd_run
res = dict(sams_run(TorchCommunicator()))
res.update(tremd_run(TorchCommunicator()))
np.savez(%(out)r + '_%%d.npz' %% dist.get_rank(), **res)
dist.barrier()
'''


def sams_run(communicator=None):
    """SAMS with the device kernel: 4 replicas (2 per rank) over 5 oscillator states, per iteration and fused."""
    from test_gpu_sams import make_sampler
    s = make_sampler(True, n_replicas=4, flatness_criteria='minimum-visits', communicator=communicator)
    s.run(12)
    s.run_fused(6)
    s.run(2)
    return dict(sams_states=np.array(s._replica_thermodynamic_states), sams_logZ=s._logZ.copy(),
                sams_u=s._energy_thermodynamic_states.copy(), sams_hist=np.array(s._state_histogram))


def tremd_run(communicator=None):
    """BASELINE configs[3] at reduced size: 8 temperatures of AlanineDipeptideVacuum (4 replicas per rank), 60 steps."""
    from openmmtools_b200 import testsystems, states, mcmc, multistate, unit
    a = testsystems.AlanineDipeptideVacuum()
    move = mcmc.LangevinSplittingDynamicsMove(timestep=2.0 * unit.femtosecond, collision_rate=5.0 / unit.picosecond, n_steps=60)
    s = multistate.ParallelTemperingSampler(mcmc_moves=move, number_of_iterations=10 ** 6, seed=77, communicator=communicator)
    s.create(states.ThermodynamicState(a.system, 300.0 * unit.kelvin), [states.SamplerState(a.positions)], storage=None,
             min_temperature=300.0 * unit.kelvin, max_temperature=600.0 * unit.kelvin, n_temperatures=8)
    s.run(4)
    return dict(pt_states=np.array(s._replica_thermodynamic_states), pt_u=s._energy_thermodynamic_states.copy())


@pytest.mark.gpu
def test_two_gpus_reproduce_single_gpu_for_sams_and_the_molecule_path(tmp_path):
    """The widened rows shard like the headline path: device SAMS (replicated jump/update kernel on the all-gathered matrix)
    and the constrained-molecule T-REMD give the same states and energies on two ranks as on one."""
    if not _two_devices():
        pytest.skip('needs 2 GPUs')
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    ref = dict(sams_run()); ref.update(tremd_run())
    script = tmp_path / 'w2.py'
    out = str(tmp_path / 'wide')
    script.write_text(WIDE_WORKER % {'root': ROOT, 'out': out})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29536', str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(2):
        d = np.load(out + '_%d.npz' % rank)
        for key, v in ref.items():
            assert np.array_equal(d[key], v), (rank, key)


SOCKET_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
from openmmtools_b200._dist import SocketCommunicator
c = SocketCommunicator()
rng = np.random.default_rng(c.rank)
shard = [(2 * c.rank + q, rng.random((7, 3)), None if q else rng.random((7, 3)), float(q), None) for q in range(2)]
g = c.gather_object(shard)
if c.rank == 0:
    assert len(g) == 2 and [it[0] for part in g for it in part] == [0, 1, 2, 3]
    r1 = np.random.default_rng(1)
    x, v = r1.random((7, 3)), r1.random((7, 3))
    assert np.array_equal(g[1][0][1], x) and np.array_equal(g[1][0][2], v) and g[1][1][2] is None and g[1][1][3] == 1.0 and g[1][0][4] is None
print('ok', c.rank)
'''


def test_socket_communicator_gathers_shards_without_pickle(tmp_path):
    """The fallback communicator (no torch.distributed): raw numeric wire format, validated sizes."""
    script = tmp_path / 's.py'
    script.write_text(SOCKET_WORKER % {'root': ROOT})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT='29561')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert sum(o.count('ok') for o in outs) == 2
    from openmmtools_b200._dist import _decode_shard
    with pytest.raises(RuntimeError):
        _decode_shard(b'\x05\x00\x00\x00' + b'\x00' * 10)
