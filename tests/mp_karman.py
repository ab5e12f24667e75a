"""Worker for the one-process-per-rank tests (torchrun): the loop of the reference's
tests/test_parallelization_von_karman.py:18-66 with this package in place of the reference modules. Every rank
checks its whole local arrays (ghost ring included) against what the reference produced on the same number of
ranks (tests/golden/karman_ranks.npz); rank 0 additionally checks the save_mpiio gather against the serial run.

GPU box:  torchrun --nproc-per-node K tests/mp_karman.py          (NCCL process group, CUDA-IPC halo)
          torchrun --nproc-per-node K tests/mp_karman.py --shared-gpu   (K processes on ONE GPU: the same CUDA-IPC peer
          stores and flag handshake between processes, kernels time-sliced; gloo process group)
CPU box:  torchrun --nproc-per-node K tests/mp_karman.py --host   (gloo; host-side logic only: the time step is
          done by the ORACLE here — tests may — while topology, bundles, CartComm.Sendrecv and save_mpiio are ours)
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    host = '--host' in sys.argv
    shared = '--shared-gpu' in sys.argv      # all ranks on ONE GPU (CUDA-IPC between processes, time-sliced kernels):
    from lattice_boltzmann_parallel_solver_b200 import dist as ldist   # NCCL refuses two ranks per device -> gloo plumbing
    from lattice_boltzmann_parallel_solver_b200 import parallelization_utils as PU
    from oracle import lbm_numpy as onp
    ldist.ensure_process_group('gloo' if host or shared else 'nccl')
    comm = ldist.comm_world()
    size, rank = comm.Get_size(), comm.Get_rank()
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'karman_ranks.npz'))
    lx, ly, d, u0, nu = 420, 180, 40, 0.1, 0.04
    omega = float(np.reciprocal(3 * nu + 0.5))
    x_size, y_size = PU.get_xy_size(size)
    cart = comm.Create_cart(dims=[x_size, y_size], periods=[True, True], reorder=False)
    c = cart.Get_coords(rank)
    nlx, nly = PU.get_local_coords(c, lx, ly, x_size, y_size)
    rho, u = onp.uniform((nlx + 2, nly + 2), 1.0, u0, 0.0)
    pc, px, py = PU.global_coord_to_local_coord(c, 3 * lx // 4, ly // 2, lx, ly, x_size, y_size)
    com = PU.communication(cart)
    trace = [u[px, py].copy()] if pc is not None else None
    if host:
        f = onp.equilibrium(rho, u)
        bc = onp.karman_parallel_bc(c, nlx, nly, lx, ly, int(x_size), int(y_size), 1.0, u0, d)
        for _ in range(11):
            f, rho, u = onp.step(f, rho, u, omega, bc, com)
            if pc is not None:
                trace.append(u[px, py].copy())
    else:
        from lattice_boltzmann_parallel_solver_b200 import boundary_utils as BU
        from lattice_boltzmann_parallel_solver_b200 import lattice_boltzmann_method as L
        f = L.equilibrium_distr_func(rho, u)
        bc = BU.parallel_von_karman_boundary_conditions(c, nlx, nly, lx, ly, x_size, y_size, 1.0, u0, d)
        for _ in range(11):
            f, rho, u = L.lattice_boltzmann_step(f, rho, u, omega, bc, com)
            if pc is not None:
                trace.append(np.array(u[px, py, ...]))
    f, rho, u = np.asarray(f), np.asarray(rho), np.asarray(u)
    assert sha(f) == str(g[f'n{size}_r{rank}_f_full']), f'rank {rank}: f differs from the reference run on {size} ranks'
    assert sha(rho) == str(g[f'n{size}_r{rank}_rho_full']), f'rank {rank}: rho'
    assert sha(u) == str(g[f'n{size}_r{rank}_u_full']), f'rank {rank}: u'
    if pc is not None:
        assert np.array_equal(np.array(trace), g[f'n{size}_probe']), 'probe trace'
    # gather through the save_mpiio replacement (tests/test_parallelization_von_karman.py:59-66)
    tmp = os.path.join(tempfile.gettempdir(), f'lbm_gather_{os.environ.get("MASTER_PORT", "0")}')
    if rank == 0:
        os.makedirs(tmp, exist_ok=True)
    comm.Barrier()
    for j in range(9):
        PU.save_mpiio(cart, os.path.join(tmp, f'f_{j}.npy'), f[1:-1, 1:-1, j])
    if rank == 0:
        F = np.stack([np.load(os.path.join(tmp, f'f_{j}.npy')) for j in range(9)], axis=-1)
        assert F.shape == (lx, ly, 9)
        assert sha(F) == str(g['serial_f11']), 'gathered populations differ from the serial run'
        print(f'OK {size} ranks ({"host/gloo" if host else ("one gpu/gloo+ipc" if shared else "gpu/nccl+ipc")})', flush=True)
    comm.Barrier()
    if not host:
        from lattice_boltzmann_parallel_solver_b200 import lattice_boltzmann_method as L
        L.release_lattices()


if __name__ == '__main__':
    main()
