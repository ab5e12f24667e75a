"""Two-row slabs + two-steps-per-pass kernel (the bench decomposition) against the single-block oracle.

    python tests/mp_slabs.py --inproc 4           one process, 4 device lattices wired with connect_blocks
                                                  (run with CUDA_DEVICE_MAX_CONNECTIONS=32: a kernel of one slab
                                                  spins on flags of kernels of other slabs enqueued later, so their
                                                  streams must not share a hardware queue)
    torchrun --nproc-per-node K tests/mp_slabs.py [--depth D] [--shared-gpu]
                                                  one process per GPU, CUDA-IPC halo (the real thing); D ghost rows and
                                                  D steps per pass; --shared-gpu: all processes on GPU 0 (gloo plumbing)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice, connect_blocks
    from oracle import lbm_c, lbm_numpy as onp
    inproc = int(sys.argv[sys.argv.index('--inproc') + 1]) if '--inproc' in sys.argv else 0
    world = int(os.environ.get('WORLD_SIZE', '1'))
    k = inproc or world
    ny, steps, omega = 512, 13, 0.8
    n = 2052 if k > 1 else 8208            # slab rows: (n + 4) x 512 >= 2^20 cells -> the bandwidth path
    nx = n * k
    rho, u = onp.sinusoidal_velocity_x((nx, ny), 0.01)
    rho = rho * np.random.default_rng(4).uniform(0.99, 1.01, rho.shape)
    f = onp.equilibrium(rho, u)
    ref = lbm_c.run(f, rho, u, omega, lbm_c.periodic(), steps)

    def padded(a, c, g=2):
        return np.ascontiguousarray(a[np.arange(c * n - g, (c + 1) * n + g) % nx])

    def check(lat, c, g=2):
        got = lat.fields(region=(g, n + g, 0, ny))
        for a, b, nm in zip(got, ref, 'f rho u'.split()):
            assert np.array_equal(a, b[c * n:(c + 1) * n]), f'slab {c}: {nm} differs from the single-block oracle'
        try:
            lat.fields(region=(0, n + 2 * g, 0, ny))
            raise SystemExit('ghost rows of a slab must not be materialisable')
        except AssertionError:
            pass

    if inproc:
        ndev = N.load().lbm_device_count()
        blocks = {(c, 0): Lattice(n + 4, ny, ghost=(2, 0), device=c % ndev) for c in range(k)}
        connect_blocks(blocks, (k, 1))
        for c in range(k):
            blocks[(c, 0)].load(padded(f, c), padded(rho, c), padded(u, c), omega)
        for lat in blocks.values():
            lat.sync()
        # All slabs share ONE GPU here, where a kernel spinning on a flag can keep the kernel that would publish it
        # from starting (a launch that changes the shared-memory carve-out waits for running kernels). So every
        # launch group is drained before the next: no kernel ever has to wait. 4 two-step passes + 5 steps = 13.
        l0 = blocks[(0, 0)].launches
        for lat in blocks.values():
            lat.set_option('fused_exact', 1)
        for chunk in (2, 2, 2, 2, 1, 1, 1, 1, 1):
            for lat in blocks.values():
                lat.run(chunk)
            for lat in blocks.values():
                lat.sync()
        assert blocks[(0, 0)].launches - l0 == 4 * 2 + 5 * 2, 'edge + interior launch per pass expected'
        for c in range(k):
            blocks[(c, 0)].sync()
            check(blocks[(c, 0)], c)
        print(f'OK {k} slabs in one process', flush=True)
    else:
        from lattice_boltzmann_parallel_solver_b200 import dist as ldist
        from lattice_boltzmann_parallel_solver_b200 import parallelization_utils as PU
        shared = '--shared-gpu' in sys.argv      # all ranks on one GPU: gloo plumbing, CUDA-IPC between the processes
        depth = int(sys.argv[sys.argv.index('--depth') + 1]) if '--depth' in sys.argv else 2
        ldist.ensure_process_group('gloo' if shared else 'nccl')
        comm = ldist.comm_world()
        rank = comm.Get_rank()
        N.set_device(0 if shared else int(os.environ.get('LOCAL_RANK', '0')))
        g = depth
        lat = Lattice(n + 2 * g, ny, ghost=(g, 0))
        lat.set_option('fused_depth', depth)
        halo = PU.communication(comm.Create_cart(dims=[k, 1], periods=[True, True]))
        halo.attach(lat)
        if '--run-host' in sys.argv:
            # the whole job in one call, pipelined over row chunks; the slab edges are finished last, in lockstep
            lat.set_option('streamed_chunk_rows', 300)
            comm.Barrier()
            pf, pr, pu = padded(f, rank, g), padded(rho, rank, g), padded(u, rank, g)
            l0 = lat.launches
            out = lat.run_host(pf, pr, pu, omega, steps)
            assert lat.launches - l0 > 2 * (n // 300), 'the pipelined schedule was not taken'
            for a, b, nm in zip(out, ref, 'f rho u'.split()):
                assert np.array_equal(a[g:n + g], b[rank * n:(rank + 1) * n]), f'slab {rank}: run_host {nm} differs from the oracle'
            comm.Barrier()
        else:
            lat.load(padded(f, rank, g), padded(rho, rank, g), padded(u, rank, g), omega)
            comm.Barrier()
            for chunk in (7, 6):          # 7: passes + a one-step launch; 6: ends on a pass (FINAL-mode materialisation)
                lat.run(chunk)
            lat.sync()
        check(lat, rank, g)
        comm.Barrier()
        if rank == 0:
            print(f'OK {k} slabs, one process per GPU' + (' (shared)' if shared else '') + f', depth {depth}' +
                  (', run_host' if '--run-host' in sys.argv else ''), flush=True)
        lat.close()


if __name__ == '__main__':
    main()
