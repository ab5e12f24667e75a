"""Workload for compute-sanitizer (run from the repository root on a GPU box):

    compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitizer_workload.py
    compute-sanitizer --tool memcheck python tools/sanitizer_workload.py

Exercises every step-kernel variant on small lattices and checks each result against the C oracle.
"""
import sys
sys.path.insert(0, '.')
import numpy as np
from lattice_boltzmann_parallel_solver_b200.engine import Lattice, connect_blocks
from oracle import lbm_numpy as onp, lbm_c as oc
# two-steps-per-pass kernel (plain and with ghost stores), mask kernel with rules, edge list kernel, materialisation
shape = (1024, 1024)
rng = np.random.default_rng(0)
rho = rng.uniform(0.9, 1.1, shape); u = rng.uniform(-0.05, 0.05, shape + (2,)); f = onp.equilibrium(rho, u)
for depth, steps in ((3, 7), (3, 6), (2, 5), (4, 8)):    # k_stepNx<3> (+ one-step launch / ending on a pass: FINAL re-run), k_step2x, k_stepNx<4>
    lat = Lattice(*shape); lat.set_option('fused_depth', depth); lat.set_option('fused_seg', 32)
    lat.probe(100, 1022, 16); lat.load(f, rho, u, 1.2); lat.run(steps)
    ref = oc.run(f, rho, u, 1.2, oc.periodic(), steps)
    print(f'{depth} steps per pass, {steps} steps ok', all(np.array_equal(a, b) for a, b in zip(lat.fields(), ref)),
          np.array_equal(lat.fields(region=(500, 503, 250, 259))[2], ref[2][500:503, 250:259]))
    lat.close()
lat = Lattice(*shape); lat.set_option('deep2', 1); lat.set_option('fused_depth', 2); lat.load(f, rho, u, 1.2); lat.run(4)
print('k_stepNx<2> ok', all(np.array_equal(a, b) for a, b in zip(lat.fields(), oc.run(f, rho, u, 1.2, oc.periodic(), 4))))
lat.close()
n = 2052
for g in (2, 3):                                          # slabs with self neighbour: edge launches with ghost stores
    blk = {(0, 0): Lattice(n + 2 * g, 512, ghost=(g, 0))}
    connect_blocks(blk, (1, 1))
    blk[(0, 0)].load_equilibrium(1.0, ux_y=0.01 * np.sin(np.arange(512) / 7.0)); blk[(0, 0)].run(7); blk[(0, 0)].sync()
    print(f'{g}-row slab (self neighbour) ran, interior rows materialised', blk[(0, 0)].fields(region=(g, g + 4, 0, 512))[1].shape)
    blk[(0, 0)].close()
# the whole job from and to host memory, pipelined over row chunks (three streams, double-buffered staging), one block and a slab
lat = Lattice(4096, 512); lat.set_option('streamed_chunk_rows', 200); lat.probe(3, 77, 64)
f5 = rng.uniform(0.9, 1.1, (4096, 512)); u5 = rng.uniform(-0.05, 0.05, (4096, 512, 2)); g5 = onp.equilibrium(f5, u5)
out = lat.run_host(g5, f5, u5, 1.2, 8)
print('run_host streamed ok', all(np.array_equal(a, b) for a, b in zip(out, oc.run(g5, f5, u5, 1.2, oc.periodic(), 8))))
lat.close()
blk = {(0, 0): Lattice(4096 + 6, 512, ghost=(3, 0))}
connect_blocks(blk, (1, 1)); blk[(0, 0)].set_option('streamed_chunk_rows', 200)
pad = lambda a: np.ascontiguousarray(a[np.arange(-3, 4099) % 4096])
out = blk[(0, 0)].run_host(pad(g5), pad(f5), pad(u5), 1.2, 8)
print('run_host on a slab (self neighbour) ok', all(np.array_equal(a[3:-3], b) for a, b in zip(out, oc.run(g5, f5, u5, 1.2, oc.periodic(), 8))))
blk[(0, 0)].close()
# cluster kernel: periodic, Couette (two cells per thread), Poiseuille (pressure-periodic stores into other CTAs' shared memory)
import lattice_boltzmann_parallel_solver_b200 as P
for name, shp, mk, scen, om in (('periodic', (100, 50), None, oc.periodic(), 1.1),
                                ('couette', (100, 100), lambda s: P.boundary_utils.couette_flow_boundary_conditions(*s, 0.05, 1.0).kind_map(s), oc.couette(0.05, 1.0), 1.0),
                                ('poiseuille', (100, 50), lambda s: P.boundary_utils.poiseuille_flow_boundary_conditions(*s, 0.3338, 0.3328).kind_map(s), oc.poiseuille(0.3338, 0.3328), 1.5)):
    r4 = rng.uniform(0.9, 1.1, shp); u4 = rng.uniform(-0.05, 0.05, shp + (2,)); f4 = onp.equilibrium(r4, u4)
    l4 = Lattice(*shp, mk(shp) if mk else None); l4.set_option('cluster', 2); l4.probe(7, 5, 64); l4.load(f4, r4, u4, om); k0 = l4.launches; l4.run(33)
    print('cluster', name, l4.launches - k0 == 1, all(np.array_equal(a, b) for a, b in zip(l4.fields(), oc.run(f4, r4, u4, om, scen, 33))))
    l4.close()
from lattice_boltzmann_parallel_solver_b200 import _native as N
# two steps per pass on a lattice WITH boundary cells: k_step2x on the clean rows, mask launches through strip windows
shape = (4096, 256)
B, BU = P.boundary_conditions, P.boundary_utils
plate = np.zeros(shape, dtype=bool); plate[1024, 100:156] = True
bundle = BU.BoundaryBundle('von_karman_serial', shape)
bundle.add(B.inlet(shape, 1.0, 0.1)).add(B.outlet()).add(B.rigid_object(plate))
rho = rng.uniform(0.9, 1.1, shape); u = rng.uniform(-0.05, 0.05, shape + (2,)); f = onp.equilibrium(rho, u)
lat = Lattice(*shape, bundle.kind_map(shape)); lat.probe(4094, 3, 16); lat.load(f, rho, u, 1.3); l0 = lat.launches; lat.run(5)
ref = oc.run(f, rho, u, 1.3, oc.karman(4096, 256, 1.0, 0.1, 56, ghost=0), 5)
print('fused with boundary strips ok', lat.launches - l0 == (1 + 3 * 2) + (1 + 2 * 2), all(np.array_equal(a, b) for a, b in zip(lat.fields(), ref)))   # a three-step and a two-step pass
lat.close()
for mode in (N.BC_MASK, N.BC_EDGE):
    lx, ly = 62, 40
    bc = P.boundary_utils.parallel_von_karman_boundary_conditions([0, 0], lx, ly, lx, ly, 1, 1, 1.0, 0.1, 8)
    r2 = rng.uniform(0.9, 1.1, (lx + 2, ly + 2)); u2 = rng.uniform(-0.05, 0.05, (lx + 2, ly + 2, 2)); f2 = onp.equilibrium(r2, u2)
    l2 = Lattice(lx + 2, ly + 2, bc.kind_map((lx + 2, ly + 2)), ghost=(1, 1), bc_mode=mode); l2.connect_self_periodic()
    l2.load(f2, r2, u2, 1.6); l2.run(70)
    ref = oc.run(f2, r2, u2, 1.6, oc.karman(lx, ly, 1.0, 0.1, 8, ghost=1), 70)
    print('karman mode', mode, all(np.array_equal(a, b) for a, b in zip(l2.fields(), ref)))
    l2.close()
bcp = P.boundary_utils.poiseuille_flow_boundary_conditions(40, 24, 0.3345, 0.3321)
r3 = rng.uniform(0.9, 1.1, (40, 24)); u3 = rng.uniform(-0.05, 0.05, (40, 24, 2)); f3 = onp.equilibrium(r3, u3)
l3 = Lattice(40, 24, bcp.kind_map((40, 24))); l3.load(f3, r3, u3, 1.5); l3.run(40)
print('poiseuille', all(np.array_equal(a, b) for a, b in zip(l3.fields(), oc.run(f3, r3, u3, 1.5, oc.poiseuille(0.3345, 0.3321), 40))))
# device history: kept velocity handles parked by the FINAL kernels writing straight into history slots
L = P.lattice_boltzmann_method
r6 = rng.uniform(0.9, 1.1, (300, 200)); u6 = rng.uniform(-0.05, 0.05, (300, 200, 2)); f6 = onp.equilibrium(r6, u6)
want, st, kept = [], (f6, r6, u6), []
a, b, c = f6, r6, u6
for _ in range(9):
    st = oc.run(*st, 1.2, oc.periodic(), 1); want.append(st[2])
    a, b, c = L.lattice_boltzmann_step(a, b, c, 1.2); kept.append(c)
np.asarray(a)
print('history', sum(h._hist is not None for h in kept) == 8, all(np.array_equal(np.asarray(h), w) for h, w in zip(kept[::-1], want[::-1])))
L.release_lattices()
