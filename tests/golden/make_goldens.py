#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_goldens.py

The reference modules are imported from where they lie (/root/reference/src); nothing is copied.
`parallelization_utils.py:3` and `boundary_utils.py:5` import `mpi4py` at module scope, which is not
installed here, so a stand-in `mpi4py.MPI` is registered in `sys.modules` first: a barrier/mailbox
communicator that runs k "ranks" as threads of this process and implements exactly the four calls the
hot path touches (Create_cart / Get_coords / Shift / Sendrecv, `parallelization_utils.py:18-49`).

Fixtures written (all float64, bit patterns of the reference's numpy results):

  kernels.npz        a2-a5: equilibrium / density / velocity / streaming on seeded random inputs
  bc_primitives.npz  a6-a11: every boundary closure applied to seeded random 10x10 (and 12x7) inputs
  topology.json      a14: get_xy_size / get_local_coords / *_in_process / global_to_local tables
  shear.npz          config 1: 100x50 periodic, omega in {0.3,1.0,1.7}; sha256 of (f,rho,u) at steps
                     1,10,100,1000 + full fields after 1000 steps at omega=1.0 + per-step max|u| trace
  couette.npz        config 2: 100x100, omega 1.0, U 0.05: digests at 1,10,100,1000 + u_x profile
                     after 1000 and 10000 steps + full fields at step 100
  poiseuille.npz     config 3: 100x50, omega 1.5, dp 0.001: same
  karman.npz         config 4: 420x180 (+ghost ring), 1-rank parallel path: probe trace (1001 samples),
                     digests at 1,2,11,100,1000 (whole arrays and interiors), a few full rows
  karman_serial.npz  milestone_6 recipe (inlet + outlet + rigid_object, no ghost ring), 200 steps
  karman_ranks.npz   k-rank runs (k = 2,4,6,8,9,14) of the parallel path, 11 steps, gathered interiors:
                     digests (all equal to the serial one) + per-rank block shapes
  ref_probe_100.npy  first 2001 samples of the cluster-generated trace
                     figures/von_karman_vortex_shedding/reynold_strouhal/vel_at_p_100.npy
  ref_vel_at_p.npy   tests/von_karman_vortex_shedding/vel_at_p.npy (12 samples)
  observables.npz    published fit numbers (figures/couette_flow/linregress.csv, figures/poiseuille_flow/*.csv),
                     the reference re-run on the published Poiseuille config (40 000 steps), viscosity fits
  ref_probe_100_full.npy  the whole 200 001-sample cluster trace at Re = 100 (Strouhal check)
"""
import hashlib
import json
import os
import sys
import threading
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------------------
# mpi4py stand-in (threads as ranks)
# --------------------------------------------------------------------------------------------------
class _World:
    def __init__(self, size):
        self.size = size
        self.barrier = threading.Barrier(size)
        self.mail = {}


class _Cart:
    def __init__(self, world, rank, dims, periods):
        self.world, self.rank = world, rank
        self.dims = [int(d) for d in dims]
        self.periods = periods

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.world.size

    def Get_coords(self, rank):
        return [rank // self.dims[1], rank % self.dims[1]]  # row-major, as MPI_Cart_create

    def _rank_of(self, c):
        return (c[0] % self.dims[0]) * self.dims[1] + (c[1] % self.dims[1])

    def Shift(self, direction, disp):
        c = self.Get_coords(self.rank)
        src = list(c)
        dst = list(c)
        src[direction] -= disp
        dst[direction] += disp
        return self._rank_of(src), self._rank_of(dst)

    def Sendrecv(self, sendbuf, dest, recvbuf=None, source=None):
        w = self.world
        w.mail[(self.rank, dest)] = np.array(sendbuf, copy=True)
        w.barrier.wait()
        recvbuf[...] = w.mail[(source, self.rank)]
        w.barrier.wait()


class _Comm:
    def __init__(self, world, rank):
        self.world, self.rank = world, rank

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.world.size

    def Create_cart(self, dims, periods, reorder=False):
        return _Cart(self.world, self.rank, dims, periods)


def _install_mpi_stub():
    mpi4py = types.ModuleType('mpi4py')
    MPI = types.ModuleType('mpi4py.MPI')
    MPI.Intracomm = _Comm
    mpi4py.MPI = MPI
    sys.modules['mpi4py'] = mpi4py
    sys.modules['mpi4py.MPI'] = MPI


_install_mpi_stub()
sys.path.insert(0, os.path.join(REF, 'src'))
import lattice_boltzmann_method as L  # noqa: E402
import boundary_conditions as B  # noqa: E402
import boundary_utils as BU  # noqa: E402
import parallelization_utils as P  # noqa: E402
import initial_values as I  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digests(tag, f, rho, u, out):
    out[tag + '_f'] = sha(f)
    out[tag + '_rho'] = sha(rho)
    out[tag + '_u'] = sha(u)


# --------------------------------------------------------------------------------------------------
def gen_kernels():
    rng = np.random.default_rng(0)
    out = {}
    for name, (nx, ny) in {'a': (17, 13), 'b': (4, 64)}.items():
        rho = rng.uniform(0.9, 1.1, (nx, ny))
        ang = rng.uniform(0, 2 * np.pi, (nx, ny))
        mag = rng.uniform(0, 0.05, (nx, ny))
        u = np.dstack([mag * np.cos(ang), mag * np.sin(ang)])
        f = rng.uniform(0.01, 0.5, (nx, ny, 9))
        rho[0, 0] = 0.0  # exercises the `where=density != 0` branch (lattice_boltzmann_method.py:126,132)
        out[name + '_rho'] = rho
        out[name + '_u'] = u
        out[name + '_f'] = f
        out[name + '_feq'] = L.equilibrium_distr_func(rho, u)
        out[name + '_density'] = L.compute_density(f)
        out[name + '_velocity'] = L.compute_velocity_field(rho, f)
        out[name + '_stream'] = L.streaming(f)
        f2, r2, u2 = L.lattice_boltzmann_step(f, rho, u, 1.3)
        out[name + '_step_f'], out[name + '_step_rho'], out[name + '_step_u'] = f2, r2, u2
    np.savez_compressed(os.path.join(OUT, 'kernels.npz'), **out)


def gen_bc_primitives():
    rng = np.random.default_rng(1)
    out = {}
    for name, (nx, ny) in {'s': (10, 10), 'r': (12, 7)}.items():
        f_pre = rng.uniform(0.01, 0.5, (nx, ny, 9))
        f_post = rng.uniform(0.01, 0.5, (nx, ny, 9))
        f_prev = rng.uniform(0.01, 0.5, (nx, ny, 9))
        rho = rng.uniform(0.9, 1.1, (nx, ny))
        u = rng.uniform(-0.05, 0.05, (nx, ny, 2))
        out[name + '_f_pre'], out[name + '_f_post'], out[name + '_f_prev'] = f_pre, f_post, f_prev
        out[name + '_rho'], out[name + '_u'] = rho, u

        def edge(which):
            m = np.zeros((nx, ny), dtype=bool)
            if which == 'x0':
                m[0, :] = True
            elif which == 'x1':
                m[-1, :] = True
            elif which == 'y0':
                m[:, 0] = True
            else:
                m[:, -1] = True
            return m

        for w in ('x0', 'x1', 'y0', 'y1'):
            out[f'{name}_rigid_{w}'] = B.rigid_wall(edge(w))(f_pre.copy(), f_post.copy())
            out[f'{name}_moving_{w}'] = B.moving_wall(edge(w), np.array([0.05, -0.02]), 1.03)(f_pre.copy(),
                                                                                              f_post.copy())
        out[name + '_inlet'] = B.inlet((nx, ny), 1.02, 0.1)(f_post.copy())
        out[name + '_outlet'] = B.outlet()(f_prev.copy(), f_post.copy())
        m = edge('x0') | edge('x1')
        out[name + '_pbc_x'] = B.periodic_with_pressure_variations(m, 0.3345, 0.3321)(f_pre.copy(), rho, u)
        m = edge('y0') | edge('y1')
        # y-variant exists in the reference (boundary_conditions.py:312-318) although nothing calls it. With an
        # all-edge mask the first branch (x) would win, so use a mask where only the y edges are equal.
        my = np.zeros((nx, ny), dtype=bool)
        my[:, 0] = True
        my[:, -1] = True
        my[0, 1] = True  # breaks boundary[0,:] == boundary[-1,:]
        # NOTE: the closure body indexes rows [0], [-2], [1], [-1] regardless of the variant
        # (boundary_conditions.py:337-344), so the y-variant only runs at all on square lattices
        # (AssertionError in equilibrium_distr_func otherwise) and then acts on ROWS with the y direction sets.
        if nx == ny:
            out[name + '_pbc_y'] = B.periodic_with_pressure_variations(my, 0.3345, 0.3321)(f_pre.copy(), rho, u)
            out[name + '_pbc_y_mask'] = my
        # plate in the middle (rigid_object mutates its mask argument: boundary_conditions.py:73-75)
        pm = np.zeros((nx, ny), dtype=bool)
        pm[nx // 4, ny // 2 - 2:ny // 2 + 2] = True
        out[name + '_plate_mask'] = pm.copy()
        out[name + '_plate'] = B.rigid_object(pm)(f_pre.copy(), f_post.copy())
        out[name + '_plate_mask_after'] = pm
        # scenario bundles applied once
        out[name + '_couette'] = BU.couette_flow_boundary_conditions(nx, ny, 0.05, 1.0)(
            f_pre.copy(), f_post.copy(), rho, u, f_prev)
        fp = f_pre.copy()
        out[name + '_poiseuille'] = BU.poiseuille_flow_boundary_conditions(nx, ny, 0.3345, 0.3321)(
            fp, f_post.copy(), rho, u, f_prev)
        out[name + '_poiseuille_fpre_after'] = fp
    np.savez_compressed(os.path.join(OUT, 'bc_primitives.npz'), **out)


def gen_topology():
    out = {'get_xy_size': {}, 'local_coords': [], 'in_process': [], 'errors': []}
    for n in range(1, 65):
        try:
            xs, ys = P.get_xy_size(n)
            out['get_xy_size'][str(n)] = [float(xs), float(ys), type(xs).__name__]
        except Exception as e:  # primes > 2 (parallelization_utils.py:75-76)
            out['get_xy_size'][str(n)] = None
            out['errors'].append([n, type(e).__name__, str(e)])
    for n in (1, 2, 4, 6, 8, 9, 14):
        xs, ys = P.get_xy_size(n)
        for lx, ly in ((420, 180), (100, 50), (37, 23)):
            for cx in range(int(xs)):
                for cy in range(int(ys)):
                    c = [cx, cy]
                    nlx, nly = P.get_local_coords(c, lx, ly, xs, ys)
                    out['local_coords'].append([n, lx, ly, cx, cy, nlx, nly])
                    for gx, gy in ((0, 0), (lx // 4, ly // 2), (lx // 4 + 1, ly // 2 + 19), (3 * lx // 4, ly // 2),
                                   (lx - 2, ly - 1), (lx - 1, 0)):
                        r = P.global_coord_to_local_coord(c, gx, gy, lx, ly, xs, ys)
                        out['in_process'].append([
                            n, lx, ly, cx, cy, gx, gy,
                            bool(P.x_in_process(c, gx, lx, xs)), bool(P.y_in_process(c, gy, ly, ys)),
                            int(P.global_to_local_direction(cx, gx, lx, xs)),
                            int(P.global_to_local_direction(cy, gy, ly, ys)),
                            None if r[0] is None else [int(r[1]), int(r[2])]])
    out['reynolds'] = float(L.reynolds_number(40, 0.1, 0.04))
    out['strouhal'] = float(L.strouhal_number(1.1308e-3, 40, 0.1))
    with open(os.path.join(OUT, 'topology.json'), 'w') as fh:
        json.dump(out, fh)


CHECK = (1, 10, 100, 1000)


def gen_shear():
    out = {}
    shape = (100, 50)
    for om in (0.3, 1.0, 1.7):
        rho, u = I.sinusoidal_velocity_x(shape, 0.01)
        if om == 1.0:
            out['rho0'], out['u0'] = rho, u
        f = L.equilibrium_distr_func(rho, u)
        amp = []
        for t in range(1, 1001):
            f, rho, u = L.lattice_boltzmann_step(f, rho, u, om)
            vmin, vmax = np.amin(u), np.amax(u)  # experiments.py:188-193
            amp.append(np.abs(vmin) if np.abs(vmin) > np.abs(vmax) else np.abs(vmax))
            if t in CHECK:
                digests(f'v_om{om}_t{t}', f, rho, u, out)
        out[f'v_om{om}_amp'] = np.array(amp)
        if om == 1.0:
            out['v_f'], out['v_rho'], out['v_u'] = f, rho, u
    # sinusoidal density (experiments.py:55-99), 50x50
    rho, u = I.sinusoidal_density_x((50, 50), 0.5, 0.08)
    out['d_rho0'], out['d_u0'] = rho, u
    f = L.equilibrium_distr_func(rho, u)
    amp = []
    for t in range(1, 1001):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 0.8)
        dmin, dmax = np.amin(rho), np.amax(rho)
        amp.append(np.abs(dmin) - 0.5 if np.abs(dmin) > np.abs(dmax) else np.abs(dmax) - 0.5)
        if t in CHECK:
            digests(f'd_t{t}', f, rho, u, out)
    out['d_amp'] = np.array(amp)
    out['d_rho'] = rho
    np.savez_compressed(os.path.join(OUT, 'shear.npz'), **out)


def gen_couette():
    out = {}
    lx, ly, om, U = 100, 100, 1.0, 0.05
    rho, u = I.density_1_velocity_0_initial((lx, ly))
    f = L.equilibrium_distr_func(rho, u)
    bc = BU.couette_flow_boundary_conditions(lx, ly, U, np.mean(rho))
    for t in range(1, 10001):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, om, bc)
        if t in CHECK:
            digests(f't{t}', f, rho, u, out)
        if t == 100:
            out['f100'], out['rho100'], out['u100'] = f, rho, u
        if t in (1000, 10000):
            out[f'ux_profile_t{t}'] = u[lx // 2, :, 0].copy()
            out[f'rho_profile_t{t}'] = rho[lx // 2, :].copy()
    digests('t10000', f, rho, u, out)
    # the published fit (figures/couette_flow/linregress.csv): 20x30, omega 1, U .05, 5000 steps
    lx, ly = 20, 30
    rho, u = I.density_1_velocity_0_initial((lx, ly))
    f = L.equilibrium_distr_func(rho, u)
    bc = BU.couette_flow_boundary_conditions(lx, ly, U, np.mean(rho))
    for t in range(5000):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.0, bc)
    out['pub_ux_profile'] = u[lx // 2, :, 0].copy()
    np.savez_compressed(os.path.join(OUT, 'couette.npz'), **out)


def gen_poiseuille():
    out = {}
    lx, ly, om, dp = 100, 50, 1.5, 0.001
    rho_in = 1 + (dp * 3) / 2
    rho_out = 1 - (dp * 3) / 2
    p_in, p_out = rho_in / 3, rho_out / 3  # experiments.py:393-398
    out['p_in'], out['p_out'] = np.float64(p_in), np.float64(p_out)
    bc = BU.poiseuille_flow_boundary_conditions(lx, ly, p_in, p_out)
    rho, u = I.density_1_velocity_0_initial((lx, ly))
    f = L.equilibrium_distr_func(rho, u)
    for t in range(1, 1001):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, om, bc)
        if t in CHECK:
            digests(f't{t}', f, rho, u, out)
        if t == 100:
            out['f100'], out['rho100'], out['u100'] = f, rho, u
    out['ux_profile_x1'] = u[1, :, 0].copy()
    out['ux_profile_mid'] = u[lx // 2, :, 0].copy()
    out['rho_centerline'] = rho[:, ly // 2].copy()
    np.savez_compressed(os.path.join(OUT, 'poiseuille.npz'), **out)


KARMAN = dict(lx=420, ly=180, d=40, u0=0.1, density_in=1.0, nu=0.04)


def _karman_rank(comm, steps, record):
    """The loop of tests/test_parallelization_von_karman.py:18-55 / experiments.py:650-704."""
    k = KARMAN
    lx, ly = k['lx'], k['ly']
    omega = np.reciprocal(3 * k['nu'] + 0.5)
    size, rank = comm.Get_size(), comm.Get_rank()
    xs, ys = P.get_xy_size(size)
    cart = comm.Create_cart(dims=[xs, ys], periods=[True, True], reorder=False)
    c = cart.Get_coords(rank)
    nlx, nly = P.get_local_coords(c, lx, ly, xs, ys)
    rho, u = I.density_1_velocity_x_u0_velocity_y_0_initial((nlx + 2, nly + 2), k['u0'])
    f = L.equilibrium_distr_func(rho, u)
    pc, px, py = P.global_coord_to_local_coord(c, 3 * lx // 4, ly // 2, lx, ly, xs, ys)
    bc = BU.parallel_von_karman_boundary_conditions(c, nlx, nly, lx, ly, xs, ys, k['density_in'], k['u0'], k['d'])
    com = P.communication(cart)
    trace = [u[px, py].copy()] if pc is not None else None
    for t in range(1, steps + 1):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, omega, bc, com)
        if pc is not None:
            trace.append(u[px, py].copy())
        record(rank, c, t, f, rho, u)
    return trace


def gen_karman():
    out = {}
    world = _World(1)
    keep = {}

    def record(rank, c, t, f, rho, u):
        if t in (1, 2, 11, 100, 1000):
            digests(f't{t}', f, rho, u, out)
            digests(f't{t}_int', f[1:-1, 1:-1], rho[1:-1, 1:-1], u[1:-1, 1:-1], out)
        if t == 11:
            keep['f11'] = f[1:-1, 1:-1].copy()
        if t == 1000:
            out['rho1000'] = rho.copy()
            out['u1000'] = u.copy()
            out['f1000_rows'] = f[[0, 1, 105, 106, 107, 211, 316, 420, 421]].copy()

    trace = _karman_rank(_Comm(world, 0), 1000, record)
    tr = np.array(trace)  # (1001, 2) components; the reference stores the norm
    out['probe_uxuy'] = tr
    out['probe_norm'] = np.array([np.linalg.norm(v) for v in tr])
    np.savez_compressed(os.path.join(OUT, 'karman.npz'), **out)
    ref12 = np.load(os.path.join(REF, 'tests/von_karman_vortex_shedding/vel_at_p.npy'))
    assert np.array_equal(out['probe_norm'][:12], ref12), 'harness does not reproduce the reference golden'
    np.save(os.path.join(OUT, 'ref_vel_at_p.npy'), ref12)
    long = np.load(os.path.join(REF, 'figures/von_karman_vortex_shedding/reynold_strouhal/vel_at_p_100.npy'))
    assert np.array_equal(out['probe_norm'], long[:1001]), 'harness does not reproduce the cluster trace'
    np.save(os.path.join(OUT, 'ref_probe_100.npy'), long[:2001])
    return keep['f11']


def gen_karman_serial(f11_parallel):
    """milestoneQuickFunctionCalls.py:293-320 — the loop that wrote the (missing) f_0..f_10.npy goldens."""
    out = {}
    k = KARMAN
    lx, ly, d, u0 = k['lx'], k['ly'], k['d'], k['u0']
    omega = np.reciprocal(3 * k['nu'] + 0.5)

    def boundary(f_pre, f_post, density=None, velocity=None, f_previous=None):
        f_post = B.inlet((lx, ly), k['density_in'], u0)(f_post)
        f_post = B.outlet()(f_previous, f_post)
        plate = np.zeros((lx, ly))
        plate[lx // 4, ly // 2 - d // 2:ly // 2 + d // 2] = 1
        return B.rigid_object(plate.astype(bool))(f_pre, f_post)

    rho, u = I.density_1_velocity_x_u0_velocity_y_0_initial((lx, ly), u0)
    f = L.equilibrium_distr_func(rho, u)
    for t in range(1, 201):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, omega, boundary)
        if t in (1, 11, 200):
            digests(f't{t}', f, rho, u, out)
        if t == 11:
            assert np.array_equal(f, f11_parallel), 'serial milestone_6 != parallel interior'
    out['u200_rows'] = u[[0, 1, 105, 106, 315, 418, 419]].copy()
    np.savez_compressed(os.path.join(OUT, 'karman_serial.npz'), **out)


def gen_karman_ranks(f11_serial):
    out = {}
    lx, ly = KARMAN['lx'], KARMAN['ly']
    for size in (2, 4, 6, 8, 9, 14):
        world = _World(size)
        blocks = {}

        def record(rank, c, t, f, rho, u):
            if t == 11:
                blocks[rank] = (list(c), f[1:-1, 1:-1].copy(), rho[1:-1, 1:-1].copy(), u[1:-1, 1:-1].copy(),
                                sha(f), sha(rho), sha(u))

        traces = [None] * size

        def run(r):
            traces[r] = _karman_rank(_Comm(world, r), 11, record)

        ts = [threading.Thread(target=run, args=(r,)) for r in range(size)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        xs, ys = (int(v) for v in P.get_xy_size(size))
        F = np.zeros((lx, ly, 9))
        shapes = []
        for r in range(size):
            c, fb, rb, ub, hf, hr, hu = blocks[r]
            x0, y0 = c[0] * (lx // xs), c[1] * (ly // ys)
            F[x0:x0 + fb.shape[0], y0:y0 + fb.shape[1]] = fb
            shapes.append([c[0], c[1], fb.shape[0], fb.shape[1]])
            out[f'n{size}_r{r}_f_full'] = hf      # whole local array incl. ghost ring
            out[f'n{size}_r{r}_rho_full'] = hr
            out[f'n{size}_r{r}_u_full'] = hu
        assert np.array_equal(F, f11_serial), f'{size}-rank gather != serial'
        out[f'n{size}_shapes'] = np.array(shapes)
        out[f'n{size}_f11'] = sha(F)
        tr = [t for t in traces if t is not None]
        assert len(tr) == 1
        out[f'n{size}_probe'] = np.array(tr[0])
    out['serial_f11'] = sha(f11_serial)
    np.savez_compressed(os.path.join(OUT, 'karman_ranks.npz'), **out)


def gen_observables():
    """Long-run observables the north star bounds to 0.5 %: published numbers (figures/*.csv), the reference run
    again here on the published configs, the viscosity fit of experiments.py:147-223 on a few omegas, and the full
    200 001-sample cluster trace at Re=100 (figures/von_karman_vortex_shedding/reynold_strouhal/vel_at_p_100.npy)."""
    import csv
    from scipy.optimize import curve_fit
    from scipy.signal import argrelextrema
    out = {}
    with open(os.path.join(REF, 'figures/couette_flow/linregress.csv')) as fh:
        row = list(csv.DictReader(fh))[0]
    out['couette_published'] = np.array([float(row['slope']), float(row['intercept']), float(row['rvalue'])])
    with open(os.path.join(REF, 'figures/poiseuille_flow/areas.csv')) as fh:
        row = list(csv.DictReader(fh))[0]
    out['poiseuille_areas_published'] = np.array([float(row['inlet']), float(row['middle']), float(row['relative_difference'])])
    out['poiseuille_fit_published'] = np.array([-4.47780943e-05, 2.64190756e-03, 1.33776993e-03])  # curve_fit.csv popt
    # Poiseuille at the published config (experiments.py:377-406): 200x60, omega 1.5, dp 0.001, 40 000 steps
    lx, ly, om, dp = 200, 60, 1.5, 0.001
    rho_in, rho_out = 1 + (dp * 3) / 2, 1 - (dp * 3) / 2
    p_in, p_out = rho_in / 3, rho_out / 3
    bc = BU.poiseuille_flow_boundary_conditions(lx, ly, p_in, p_out)
    rho, u = I.density_1_velocity_0_initial((lx, ly))
    f = L.equilibrium_distr_func(rho, u)
    for t in range(40000):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, om, bc)
    out['poiseuille_p'] = np.array([p_in, p_out])
    out['poiseuille_ux_x1'] = u[1, :, 0].copy()
    out['poiseuille_ux_mid'] = u[lx // 2, :, 0].copy()
    out['poiseuille_rho_centerline'] = rho[:, ly // 2].copy()
    digests('poiseuille_t40000', f, rho, u, out)
    # viscosity vs omega (experiments.py:147-223), 50x50, 2500 steps, a subset of the 50 omegas
    shape, steps = (50, 50), 2500
    omegas = np.linspace(0.01, 1.99, 50)[[2, 12, 24, 37, 47]]
    out['visc_omegas'] = omegas
    sims = [[], []]
    for i, initial in enumerate([I.sinusoidal_density_x(shape, 0.5, 0.08), I.sinusoidal_velocity_x(shape, 0.08)]):
        for k, omg in enumerate(omegas):
            rho, u = initial
            f = L.equilibrium_distr_func(rho, u)
            amp = []
            for _ in range(steps):
                f, rho, u = L.lattice_boltzmann_step(f, rho, u, omg)
                if i == 0:
                    lo, hi = np.amin(rho), np.amax(rho)
                    amp.append(np.abs(lo) - 0.5 if np.abs(lo) > np.abs(hi) else np.abs(hi) - 0.5)
                else:
                    lo, hi = np.amin(u), np.amax(u)
                    amp.append(np.abs(lo) if np.abs(lo) > np.abs(hi) else np.abs(hi))
            amp = np.array(amp)
            out[f'visc_amp_{i}_{k}'] = amp
            if i == 0:
                idx = argrelextrema(amp, np.greater)
                v = curve_fit(lambda t, v: 0.08 * np.exp(-v * np.power(2 * np.pi / shape[0], 2) * t),
                              np.array(idx).squeeze(), amp[idx])[0][0]
            else:
                v = curve_fit(lambda t, v: 0.08 * np.exp(-v * np.power(2 * np.pi / shape[-1], 2) * t),
                              np.arange(0, steps), amp)[0][0]
            sims[i].append(v)
    out['visc_sim_density'] = np.array(sims[0])
    out['visc_sim_velocity'] = np.array(sims[1])
    out['visc_true'] = (1 / 3) * (1 / omegas - 0.5)
    np.savez_compressed(os.path.join(OUT, 'observables.npz'), **out)
    trace = np.load(os.path.join(REF, 'figures/von_karman_vortex_shedding/reynold_strouhal/vel_at_p_100.npy'))
    np.save(os.path.join(OUT, 'ref_probe_100_full.npy'), trace)


if __name__ == '__main__':
    which = sys.argv[1:] or ['kernels', 'bc', 'topology', 'shear', 'couette', 'poiseuille', 'karman']
    if 'kernels' in which:
        gen_kernels()
    if 'bc' in which:
        gen_bc_primitives()
    if 'topology' in which:
        gen_topology()
    if 'shear' in which:
        gen_shear()
    if 'couette' in which:
        gen_couette()
    if 'poiseuille' in which:
        gen_poiseuille()
    if 'observables' in which:
        gen_observables()
    if 'karman' in which:
        f11 = gen_karman()
        gen_karman_serial(f11)
        gen_karman_ranks(f11)
    print('goldens written to', OUT)
