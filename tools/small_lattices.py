"""us per step of the launch-bound lattices (configs 1-4 of BASELINE.json and a few mid sizes): CUDA-graph replay of 32
captured steps, with and without programmatic dependent launch between the step kernels. Run on a GPU box:

    python tools/small_lattices.py
"""
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import lattice_boltzmann_parallel_solver_b200 as P
from lattice_boltzmann_parallel_solver_b200.engine import Lattice
from oracle import lbm_numpy as onp

BU = P.boundary_utils


def timeit(name, lat, n):
    st = torch.cuda.ExternalStream(lat.stream)
    # cluster kernel (whole lattice in distributed shared memory, all n steps in one launch) where the lattice fits
    lat.set_option('cluster', 2)
    lat.run(200)
    lat.sync()
    l0 = lat.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    lat.run(n)
    e1.record(st)
    lat.sync()
    if lat.launches - l0 < n // 64:
        print(f'{name:34s} {n} steps: {1e3 * e0.elapsed_time(e1) / n:6.2f} us/step in {lat.launches - l0} cluster launch(es)', flush=True)
    lat.set_option('cluster', 0)
    res = []
    for pdl in (1, 0):
        lat.set_option('pdl', pdl)
        lat.run(200)
        lat.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        lat.run(n)
        e1.record(st)
        lat.sync()
        t2 = time.perf_counter()
        res.append((1e6 * (t2 - t0) / n, 1e3 * e0.elapsed_time(e1) / n))
    lat.set_option('pdl', 1)
    lat.set_option('cluster', 1)      # default: the first four calls are timed alternately on both paths, the faster one is kept
    for _ in range(6):
        lat.run(256)
    lat.sync()
    lat.run(256)                      # (the samples are read back without blocking: resolved at the first call after they completed)
    lat.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    lat.run(n)
    e1.record(st)
    lat.sync()
    print(f'{name:34s} auto (library default): {1e3 * e0.elapsed_time(e1) / n:6.2f} us/step', flush=True)
    (w1, g1), (w0, g0) = res
    print(f'{name:34s} graph replay: {g1:6.2f} us/step (wall {w1:6.2f}) = {lat.nx * lat.ny / g1:9.1f} MLUPS | '
          f'without PDL {g0:6.2f} us/step (wall {w0:6.2f})', flush=True)


rho, u = onp.sinusoidal_velocity_x((100, 50), 0.01)
f = onp.equilibrium(rho, u)
lat = Lattice(100, 50); lat.load(f, rho, u, 1.0); timeit('config1 shear 100x50', lat, 20000); lat.close()
rho, u = onp.uniform((100, 100)); f = onp.equilibrium(rho, u)
lat = Lattice(100, 100, BU.couette_flow_boundary_conditions(100, 100, 0.05, 1.0).kind_map((100, 100)))
lat.load(f, rho, u, 1.0); timeit('config2 couette 100x100', lat, 20000); lat.close()
rho, u = onp.uniform((100, 50)); f = onp.equilibrium(rho, u)
lat = Lattice(100, 50, BU.poiseuille_flow_boundary_conditions(100, 50, 0.3338, 0.3328).kind_map((100, 50)))
lat.load(f, rho, u, 1.5); timeit('config3 poiseuille 100x50', lat, 20000); lat.close()
rho, u = onp.uniform((422, 182), 1.0, 0.1, 0.0); f = onp.equilibrium(rho, u)
bc = BU.parallel_von_karman_boundary_conditions([0, 0], 420, 180, 420, 180, 1, 1, 1.0, 0.1, 40)
lat = Lattice(422, 182, bc.kind_map((422, 182)), ghost=(1, 1)); lat.connect_self_periodic(); lat.load(f, rho, u, 1.6)
timeit('config4 karman 422x182 (ghost)', lat, 20000)
lat.probe(316, 91, 65536); timeit('config4 karman + probe', lat, 20000); lat.close()
rho, u = onp.uniform((420, 180), 1.0, 0.1, 0.0); f = onp.equilibrium(rho, u)
lat = Lattice(420, 180); lat.load(f, rho, u, 1.6); timeit('periodic 420x180', lat, 20000); lat.close()
for n in (256, 512, 1000):
    lat = Lattice(n, n); lat.load_equilibrium(1.0, ux_y=0.01 * np.sin(np.arange(n) / 9.0)); timeit(f'periodic {n}x{n}', lat, 5000); lat.close()
