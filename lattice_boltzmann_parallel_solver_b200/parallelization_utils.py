"""Drop-in for the reference's `src/parallelization_utils.py`: same names and arguments.

`communication(comm)` no longer moves faces through four blocking `Sendrecv` calls per step
(reference: src/parallelization_utils.py:34-49). It returns a `HaloExchange` that, handed to
`lattice_boltzmann_step`, wires the device lattices of the ranks together ONCE (CUDA-IPC handles exchanged
through the process group) so that the fused step kernel stores ghost cells straight into the neighbour's HBM
over NVLink. Called directly on a host array it still performs the reference's four exchanges through
`comm.Sendrecv` (pure data movement — used by host-side tests).
"""
from typing import Callable, Tuple

import numpy as np

from . import _native as N


class HaloExchange:
    """What `communication(comm)` returns. `comm` is a Cartesian communicator with the mpi4py method names the
    reference uses (this package's `dist.CartComm`, or a real `mpi4py` Cartcomm)."""

    def __init__(self, comm):
        self.comm = comm
        # neighbour ranks, precomputed like the reference's closure does (:18-21)
        self.left_src, self.left_dst = comm.Shift(direction=0, disp=-1)
        self.right_src, self.right_dst = comm.Shift(direction=0, disp=1)
        self.bottom_src, self.bottom_dst = comm.Shift(direction=1, disp=-1)
        self.top_src, self.top_dst = comm.Shift(direction=1, disp=1)

    # -- host arrays: the reference's semantics, verbatim order (x faces first, so y faces carry the corners) --
    def __call__(self, f: np.ndarray) -> np.ndarray:
        c = self.comm
        buf = np.empty_like(f[-1, ...])
        c.Sendrecv(np.ascontiguousarray(f[1, ...]), self.left_dst, recvbuf=buf, source=self.left_src)
        f[-1, ...] = buf
        buf = np.empty_like(f[0, ...])
        c.Sendrecv(np.ascontiguousarray(f[-2, ...]), self.right_dst, recvbuf=buf, source=self.right_src)
        f[0, ...] = buf
        buf = np.empty_like(np.ascontiguousarray(f[:, -1, :]))
        c.Sendrecv(np.ascontiguousarray(f[:, 1, :]), self.bottom_dst, recvbuf=buf, source=self.bottom_src)
        f[:, -1, :] = buf
        buf = np.empty_like(np.ascontiguousarray(f[:, 0, :]))
        c.Sendrecv(np.ascontiguousarray(f[:, -2, :]), self.top_dst, recvbuf=buf, source=self.top_src)
        f[:, 0, :] = buf
        return f

    # -- device lattices --------------------------------------------------------------------------------------
    def _rank_of(self, dx, dy):
        c = self.comm
        me = c.Get_coords(c.Get_rank())
        dims = [int(d) for d in c.dims]
        return ((me[0] + dx) % dims[0]) * dims[1] + ((me[1] + dy) % dims[1])

    def attach(self, lattice):
        """Connect the eight ghost-ring neighbours of `lattice` (collective over the communicator)."""
        mine = lattice.halo_export()
        everyone = self.comm.allgather(bytes(mine))
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                if (dx == 0 and dy == 0) or (dx and not lattice.ghost[0]) or (dy and not lattice.ghost[1]):
                    continue
                peer = N.HaloExport.from_buffer_copy(everyone[self._rank_of(dx, dy)])
                lattice.halo_connect((dx + 1) * 3 + dy + 1, peer)
        lattice.halo_finalize()
        self._barrier()

    def before_load(self, lattice):
        """A fresh upload stores ghost cells into the neighbours' buffers: every rank must be done reading the
        previous run's state first."""
        self._barrier()

    def after_load(self, lattice):
        """Every rank's first-collision ghost stores must have landed before anyone takes the first step."""
        self._barrier()

    def _barrier(self):
        bar = getattr(self.comm, 'Barrier', None) or getattr(self.comm, 'barrier', None)
        if bar is not None:
            bar()


def communication(comm) -> Callable[[np.ndarray], np.ndarray]:
    """The communication step for the parallel implementation (reference: src/parallelization_utils.py:6-52)."""
    return HaloExchange(comm)


def get_xy_size(total_number_of_procces: int) -> Tuple[int, int]:
    """Most-square process grid with x_size <= y_size; primes above 2 are rejected
    (reference: src/parallelization_utils.py:55-92 — including its return types: ints for one process,
    numpy floats otherwise, which then flow into Create_cart and get_local_coords)."""
    n = total_number_of_procces
    if n > 2 and all(n % k for k in range(2, n)):
        raise Exception('This implementation does not work if number of nodes is a prime (excluding 1 and 2)')
    if n <= 1:
        return 1, 1
    lower = upper = np.ceil(np.sqrt(n))
    while lower * upper != n:
        if lower * upper > n:
            lower -= 1
        else:
            upper += 1
    return lower, upper


def get_local_coords(coords2d: list, lx: int, ly: int, x_size: int, y_size: int) -> Tuple[int, int]:
    """Block size of a rank: L // P per direction, the last rank takes the remainder
    (reference: src/parallelization_utils.py:95-120)."""
    n_local_x, n_local_y = lx // x_size, ly // y_size
    if coords2d[0] + 1 == x_size:
        n_local_x = lx - n_local_x * (x_size - 1)
    if coords2d[1] + 1 == y_size:
        n_local_y = ly - n_local_y * (y_size - 1)
    return int(n_local_x), int(n_local_y)


def global_to_local_direction(coord1d: int, global_dir: int, lattice_dir: int, dir_size: int):
    """Global index -> index in the ghost-padded local array (reference: src/parallelization_utils.py:123-137)."""
    return int(global_dir - coord1d * (lattice_dir // dir_size)) + 1


def x_in_process(coord2d: list, x_coord: int, lx: int, processes_in_x: int) -> bool:
    """Does this rank own global column x_coord? (reference: src/parallelization_utils.py:165-181)"""
    first = coord2d[0] * (lx // processes_in_x)
    last = lx - 1 if coord2d[0] == processes_in_x - 1 else (coord2d[0] + 1) * (lx // processes_in_x) - 1
    return first <= x_coord <= last


def y_in_process(coord2d: list, y_coord: int, ly: int, processes_in_y: int) -> bool:
    """Does this rank own global row y_coord? (reference: src/parallelization_utils.py:184-200)"""
    first = coord2d[1] * (ly // processes_in_y)
    last = ly - 1 if coord2d[1] == processes_in_y - 1 else (coord2d[1] + 1) * (ly // processes_in_y) - 1
    return first <= y_coord <= last


def global_coord_to_local_coord(coord2d: list, global_x: int, global_y: int, lx: int, ly: int, x_size: int,
                                y_size: int) -> Tuple[int, int]:
    """(coord2d, local_x, local_y) if this rank owns the cell, else (None, None, None)
    (reference: src/parallelization_utils.py:140-162)."""
    if x_in_process(coord2d, global_x, lx, x_size) and y_in_process(coord2d, global_y, ly, y_size):
        return (coord2d, global_to_local_direction(coord2d[0], global_x, lx, x_size),
                global_to_local_direction(coord2d[1], global_y, ly, y_size))
    return None, None, None


def save_mpiio(comm, fn: str, g_kl: np.ndarray):
    """Write the global 2-D array whose blocks live on the ranks of a Cartesian communicator to ONE .npy file
    readable by numpy.load (reference: src/parallelization_utils.py:203-253, which uses MPI-IO and, on current
    numpy, fails at `np.asscalar`). Here: blocks are gathered through the process group and rank 0 writes a
    standard version-1.0 .npy — same file contents. Output path, not the time step."""
    g_kl = np.ascontiguousarray(np.asarray(g_kl))
    assert g_kl.ndim == 2
    rank = comm.Get_rank()
    coords = comm.Get_coords(rank)
    pieces = comm.allgather((coords, g_kl))
    if rank == 0:
        dims = [int(d) for d in comm.dims]
        rows = [np.concatenate([blk for c, blk in sorted(pieces, key=lambda p: p[0]) if c[0] == i], axis=1)
                for i in range(dims[0])]
        np.save(fn if str(fn).endswith('.npy') else str(fn) + '.npy', np.concatenate(rows, axis=0))
    bar = getattr(comm, 'Barrier', None)
    if bar is not None:
        bar()
