"""The reference's x_strouhal loop shape (src/experiments.py:699-704) through the drop-in modules: one
lattice_boltzmann_step per iteration and ONE velocity cell read after it — wall-clock us per iteration on the von Karman
420 x 180 lattice (parallel path on one rank, ghost ring), and the same loop without the read (deferred batches).

    python tools/per_step_read.py [steps]
"""
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import lattice_boltzmann_parallel_solver_b200 as P
from lattice_boltzmann_parallel_solver_b200.dist import comm_world

L, BU, PU = P.lattice_boltzmann_method, P.boundary_utils, P.parallelization_utils
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
lx, ly, d = 420, 180, 40
comm = comm_world().Create_cart((1, 1), periods=(True, True))
for read in (True, False):
    density = np.ones((lx + 2, ly + 2))
    velocity = np.zeros((lx + 2, ly + 2, 2)); velocity[..., 0] = 0.1
    f = L.equilibrium_distr_func(density, velocity)
    bc = BU.parallel_von_karman_boundary_conditions([0, 0], lx, ly, lx, ly, 1, 1, 1.0, 0.1, d)
    com = PU.communication(comm)
    px, py = 3 * lx // 4 + 1, ly // 2 + 1
    trace = []
    for i in range(steps + 200):
        if i == 200:
            np.asarray(velocity[px, py]); t0 = time.perf_counter()
        f, density, velocity = L.lattice_boltzmann_step(f, density, velocity, 1.6, bc, com)
        if read:
            trace.append(np.linalg.norm(velocity[px, py, ...]))
    np.asarray(velocity[px, py])
    dt = time.perf_counter() - t0
    print(f"{'one cell read after every step' if read else 'no reads (deferred batches)   '}: {1e6 * dt / steps:7.2f} us per iteration "
          f"({steps} steps, {lx * ly * steps / dt / 1e6:8.1f} MLUPS)", flush=True)
    L.release_lattices()
