"""Minimal host-side plumbing for multi-process runs (one process per GPU).

Only small control messages cross here (the 128-byte NCCL unique id, a seed); all bulk traffic is NCCL inside the
engine.  ``torch.distributed`` is used when it is already initialised (torchrun), otherwise a tiny TCP
exchange on MASTER_ADDR:MASTER_PORT+1.  Replaces the role of mpiplus' bcast
(/root/reference/openmmtools/multistate/replicaexchange.py:255).
"""
import os
import socket
import time


class TorchCommunicator:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank()
        self.world_size = dist.get_world_size()

    def bcast_bytes(self, data, n):
        import torch
        t = torch.zeros(n, dtype=torch.uint8)
        if self.rank == 0:
            t = torch.tensor(list(data[:n].ljust(n, b'\0')), dtype=torch.uint8)
        if self.dist.get_backend() == 'nccl':
            t = t.cuda()
        self.dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    def barrier(self):
        self.dist.barrier()

    def gather_object(self, obj):
        """List of every rank's object on rank 0, None elsewhere (checkpoint gather; small control-plane traffic)."""
        out = [None] * self.world_size if self.rank == 0 else None
        self.dist.gather_object(obj, out, dst=0)
        return out


class SocketCommunicator:
    """Rank 0 listens; the others connect and receive.  Enough for a broadcast of a few bytes."""

    def __init__(self):
        self.rank = int(os.environ.get('RANK', '0'))
        self.world_size = int(os.environ.get('WORLD_SIZE', '1'))
        self.addr = os.environ.get('MASTER_ADDR', '127.0.0.1')
        self.port = int(os.environ.get('MASTER_PORT', '29500')) + 1

    def bcast_bytes(self, data, n):
        if self.world_size == 1:
            return data
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((self.addr, self.port))
            srv.listen(self.world_size)
            for _ in range(self.world_size - 1):
                c, _a = srv.accept()
                c.sendall(data[:n].ljust(n, b'\0'))
                c.close()
            srv.close()
            return data[:n].ljust(n, b'\0')
        for _ in range(600):
            try:
                c = socket.create_connection((self.addr, self.port), timeout=5)
                break
            except OSError:
                time.sleep(0.1)
        else:
            raise RuntimeError('could not reach rank 0')
        buf = b''
        while len(buf) < n:
            chunk = c.recv(n - len(buf))
            if not chunk:
                break
            buf += chunk
        c.close()
        return buf

    def barrier(self):
        self.bcast_bytes(b'x', 1)

    def gather_object(self, obj):
        import pickle
        import struct
        if self.world_size == 1:
            return [obj]
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((self.addr, self.port + 1))
            srv.listen(self.world_size)
            out = [None] * self.world_size
            out[0] = obj
            for _ in range(self.world_size - 1):
                c, _a = srv.accept()
                hdr = b''
                while len(hdr) < 12:
                    hdr += c.recv(12 - len(hdr))
                r, n = struct.unpack('<iq', hdr)
                buf = b''
                while len(buf) < n:
                    buf += c.recv(min(1 << 20, n - len(buf)))
                out[r] = pickle.loads(buf)
                c.close()
            srv.close()
            return out
        payload = pickle.dumps(obj)
        for _ in range(600):
            try:
                c = socket.create_connection((self.addr, self.port + 1), timeout=5)
                break
            except OSError:
                time.sleep(0.1)
        else:
            raise RuntimeError('could not reach rank 0')
        c.sendall(struct.pack('<iq', self.rank, len(payload)) + payload)
        c.close()
        return None


def default_communicator():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return TorchCommunicator()
    except ImportError:
        pass
    return SocketCommunicator()
