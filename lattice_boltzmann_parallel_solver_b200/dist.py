"""One-process-per-GPU plumbing: the slice of the mpi4py communicator API the reference's hot path touches
(`Create_cart / Get_coords / Shift / Sendrecv`, src/parallelization_utils.py:18-49; `Get_rank / Get_size`,
src/experiments.py:606-614), implemented over `torch.distributed` (NCCL on the GPU box, gloo in CPU tests) or,
with a single process, over nothing at all.

This is the process-group side only. The per-step ghost exchange itself is NOT done through these calls on the
device path: `parallelization_utils.communication` wires CUDA-IPC peer mappings once and the step kernel stores
ghost cells directly into the neighbour's memory (include/lbm_b200.h, "Halo exchange").
"""
import os

import numpy as np


def _dist():
    import torch.distributed as dist
    return dist


def ensure_process_group(backend=None):
    """Initialise torch.distributed from the torchrun environment if this is a multi-process launch."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1:
        return False
    dist = _dist()
    if not dist.is_initialized():
        import torch
        if backend is None:   # LBM_DIST_BACKEND=gloo: several ranks on ONE GPU (NCCL refuses that; the halo itself is CUDA-IPC)
            backend = os.environ.get('LBM_DIST_BACKEND') or ('nccl' if torch.cuda.is_available() else 'gloo')
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if backend == 'nccl':
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend=backend)
    return True


def _flush_queued_steps():
    # a rank must not block on the host while steps of a lattice with halo neighbours are still queued
    from . import lattice_boltzmann_method as lbm
    lbm.flush_all()


class WorldComm:
    """Stands where `MPI.COMM_WORLD` stands in the reference's drivers."""

    def __init__(self, group=None):
        self.group = group
        self._multi = ensure_process_group() if group is None else True

    def Get_rank(self):
        return _dist().get_rank(self.group) if self._multi else 0

    def Get_size(self):
        return _dist().get_world_size(self.group) if self._multi else 1

    def Barrier(self):
        if self._multi:
            _flush_queued_steps()
            _dist().barrier(self.group)

    barrier = Barrier

    def allgather(self, obj):
        if not self._multi:
            return [obj]
        _flush_queued_steps()
        out = [None] * self.Get_size()
        _dist().all_gather_object(out, obj, group=self.group)
        return out

    def Create_cart(self, dims, periods=(True, True), reorder=False):
        return CartComm(self, dims, periods)


class CartComm(WorldComm):
    """Periodic 2-D Cartesian topology, row-major rank order (what MPI_Cart_create gives with reorder=False;
    reference: src/experiments.py:613, dims = get_xy_size(size) — possibly numpy floats)."""

    def __init__(self, world, dims, periods=(True, True)):
        self.group = world.group
        self._multi = world._multi
        self.dims = [int(d) for d in dims]
        self.periods = [bool(p) for p in periods]
        assert len(self.dims) == 2 and all(self.periods), 'the reference only builds fully periodic 2-D topologies'
        assert self.dims[0] * self.dims[1] == self.Get_size(), \
            f'topology {self.dims} does not match {self.Get_size()} processes'

    def Get_coords(self, rank):
        return [int(rank) // self.dims[1], int(rank) % self.dims[1]]

    def Get_cart_rank(self, coords):
        return (int(coords[0]) % self.dims[0]) * self.dims[1] + (int(coords[1]) % self.dims[1])

    def Shift(self, direction, disp):
        """(source, dest) ranks for a shift by `disp` along `direction`, as MPI_Cart_shift."""
        me = self.Get_coords(self.Get_rank())
        src, dst = list(me), list(me)
        src[direction] -= disp
        dst[direction] += disp
        return self.Get_cart_rank(src), self.Get_cart_rank(dst)

    def neighbour(self, dx, dy):
        me = self.Get_coords(self.Get_rank())
        return self.Get_cart_rank([me[0] + dx, me[1] + dy])

    def Sendrecv(self, sendbuf, dest, recvbuf=None, source=None, **_):
        """Blocking exchange of host arrays (numpy). Used by `communicate(f)` when it is called on host arrays."""
        rank = self.Get_rank()
        sendbuf = np.ascontiguousarray(sendbuf)
        if dest == rank and source == rank:
            recvbuf[...] = sendbuf
            return
        import torch
        _flush_queued_steps()
        dist = _dist()
        on_gpu = dist.get_backend(self.group) == 'nccl'
        dev = torch.device('cuda', torch.cuda.current_device()) if on_gpu else torch.device('cpu')
        s = torch.from_numpy(sendbuf).to(dev)
        r = torch.empty(recvbuf.shape, dtype=s.dtype, device=dev)
        ops = [dist.P2POp(dist.isend, s, dest, self.group), dist.P2POp(dist.irecv, r, source, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        recvbuf[...] = r.cpu().numpy()


COMM_WORLD = None


def comm_world():
    global COMM_WORLD
    if COMM_WORLD is None:
        COMM_WORLD = WorldComm()
    return COMM_WORLD
