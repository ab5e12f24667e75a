"""The von Karman rule set (inlet row, outlet rows, plate) on slabs with D ghost rows: D = 3 (or 2) steps per pass on every
rank — the fluid two-step kernel on rows whose cone is all fluid, strip windows next to boundary rows and next to the
slab edges (those launches read the ghost rows, store into the neighbours' and carry the flag handshake) — against the
single-block C oracle on a field that varies along both axes.

    torchrun --nproc-per-node K tests/mp_karman_slabs.py [--depth D] [--shared-gpu]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200 import dist as ldist
    from lattice_boltzmann_parallel_solver_b200 import parallelization_utils as PU
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    from oracle import lbm_c, lbm_numpy as onp
    shared = '--shared-gpu' in sys.argv
    ldist.ensure_process_group('gloo' if shared else 'nccl')
    comm = ldist.comm_world()
    rank, k = comm.Get_rank(), comm.Get_size()
    N.set_device(0 if shared else int(os.environ.get('LOCAL_RANK', '0')))
    depth = int(sys.argv[sys.argv.index('--depth') + 1]) if '--depth' in sys.argv else 3
    ny, n, g, steps = 512, 2052, depth, 13
    nxg = n * k
    omega = float(np.reciprocal(3 * 0.04 + 0.5))
    d = int(ny / 4.5) // 2 * 2
    rng = np.random.default_rng(17)
    rho = rng.uniform(0.98, 1.02, (nxg, ny))
    u = np.zeros((nxg, ny, 2))
    u[..., 0] = 0.1 * rng.uniform(0.9, 1.1, (nxg, ny))
    u[..., 1] = 0.01 * rng.uniform(-1, 1, (nxg, ny))
    f = onp.equilibrium(rho, u)
    ref = lbm_c.run(f, rho, u, omega, lbm_c.karman(nxg, ny, 1.0, 0.1, d, ghost=0), steps)

    def padded(a):
        return np.ascontiguousarray(a[np.arange(rank * n - g, (rank + 1) * n + g) % nxg])

    km = bench.karman_slab_kind_map(nxg, ny, rank * n - g, n + 2 * g)
    lat = Lattice(n + 2 * g, ny, km, ghost=(g, 0))
    PU.communication(comm.Create_cart(dims=[k, 1], periods=[True, True])).attach(lat)
    lat.load(padded(f), padded(rho), padded(u), omega)
    comm.Barrier()
    l0 = lat.launches
    for chunk in (7, 6):               # 3 passes + a one-step launch; 3 passes: the job ENDS on a pass (materialised from the windows)
        lat.run(chunk)
    lat.sync()
    launches = lat.launches - l0
    got = lat.fields(region=(g, n + g, 0, ny))
    for a, b, nm in zip(got, ref, 'f rho u'.split()):
        b = b[rank * n:(rank + 1) * n]
        assert np.array_equal(a, b), (f'rank {rank}: {nm} differs from the single-block oracle in '
                                      f'{int(np.count_nonzero(a != b))} values, rows {sorted(set(np.argwhere(a != b)[:, 0].tolist()))[:8]}')
    everyone = comm.allgather((rank, launches, km.is_trivial))
    comm.Barrier()
    if rank == 0:
        # 13 one-step launches would be >= 13 (x2 with the fix-up kernel); two-step passes: 6 passes + 1 single step
        print(f'OK {k} karman slabs, {depth} steps per pass' + (' (shared)' if shared else '') + f', launches per rank {[(r, l, "fluid" if t else "bc") for r, l, t in everyone]}',
              flush=True)
    lat.close()


if __name__ == '__main__':
    main()
