"""Pins the oracle (oracle/lbm_numpy.py and oracle/lbm_oracle.c) bit-for-bit against fixtures produced by the
unmodified reference (tests/golden/make_goldens.py). CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import lbm_c as oc
from oracle import lbm_numpy as onp
from tests.helpers import sha

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def check_digests(g, tag, f, rho, u):
    assert sha(f) == str(g[tag + '_f']), tag + ' f'
    assert sha(rho) == str(g[tag + '_rho']), tag + ' rho'
    assert sha(u) == str(g[tag + '_u']), tag + ' u'


@pytest.mark.parametrize('n', ['a', 'b'])
def test_kernels_numpy_and_c(n):
    g = load('kernels.npz')
    rho, u, f = g[n + '_rho'], g[n + '_u'], g[n + '_f']
    assert np.array_equal(onp.equilibrium(rho, u), g[n + '_feq'])
    assert np.array_equal(oc.equilibrium(rho, u), g[n + '_feq'])
    assert np.array_equal(onp.density(f), g[n + '_density'])
    assert np.array_equal(onp.velocity(rho, f), g[n + '_velocity'])
    assert np.array_equal(onp.stream(f), g[n + '_stream'])
    for impl in ('np', 'c'):
        if impl == 'np':
            f2, r2, u2 = onp.step(f, rho, u, 1.3)
        else:
            f2, r2, u2 = oc.run(f, rho, u, 1.3, oc.periodic(), 1)
        assert np.array_equal(f2, g[n + '_step_f'])
        assert np.array_equal(r2, g[n + '_step_rho'])
        assert np.array_equal(u2, g[n + '_step_u'])


@pytest.mark.parametrize('n', ['s', 'r'])
def test_bc_primitives_numpy(n):
    g = load('bc_primitives.npz')
    f_pre, f_post, f_prev, rho, u = (g[n + k] for k in ('_f_pre', '_f_post', '_f_prev', '_rho', '_u'))
    nx, ny = rho.shape
    for w in ('x0', 'x1', 'y0', 'y1'):
        m = onp._edge((nx, ny), w)
        assert np.array_equal(onp.rigid_wall(m)(f_pre.copy(), f_post.copy()), g[f'{n}_rigid_{w}'])
        assert np.array_equal(onp.moving_wall(m, np.array([0.05, -0.02]), 1.03)(f_pre.copy(), f_post.copy()),
                              g[f'{n}_moving_{w}'])
    assert np.array_equal(onp.inlet((nx, ny), 1.02, 0.1)(f_post.copy()), g[n + '_inlet'])
    assert np.array_equal(onp.outlet()(f_prev.copy(), f_post.copy()), g[n + '_outlet'])
    assert np.array_equal(onp.pbc_pressure_x(0.3345, 0.3321)(f_pre.copy(), rho, u), g[n + '_pbc_x'])
    pm = g[n + '_plate_mask'].copy()
    assert np.array_equal(onp.rigid_object(pm)(f_pre.copy(), f_post.copy()), g[n + '_plate'])
    assert np.array_equal(pm, g[n + '_plate_mask_after'])
    assert np.array_equal(onp.couette_bc(nx, ny, 0.05, 1.0)(f_pre.copy(), f_post.copy(), rho, u, f_prev),
                          g[n + '_couette'])
    fp = f_pre.copy()
    assert np.array_equal(onp.poiseuille_bc(nx, ny, 0.3345, 0.3321)(fp, f_post.copy(), rho, u, f_prev),
                          g[n + '_poiseuille'])
    assert np.array_equal(fp, g[n + '_poiseuille_fpre_after'])


def test_reference_unit_test_constants():
    """The known answers of the reference's tests/test_boundary_conditions.py:87-121 (10x10, p_in=1, p_out=0.1)."""
    shape = (10, 10)
    f = np.ones(shape + (9,))
    rho = np.ones(shape)
    u = np.zeros(shape + (2,))
    u[..., 0] = 0.1
    out = onp.pbc_pressure_x(1, 0.1)(f, rho, u)
    assert np.allclose(out[0, :, 1], 1.29555)
    assert np.allclose(out[0, :, 5], 1.073888) and np.allclose(out[0, :, 8], 1.073888)
    assert np.allclose(out[-1, :, 3], 0.943222)
    assert np.allclose(out[-1, :, 6], 0.9858) and np.allclose(out[-1, :, 7], 0.9858)
    # moving wall, tests/test_boundary_conditions.py:62-85
    m = onp._edge(shape, 'y1')
    out = onp.moving_wall(m, np.array([2, 0]), 1)(np.ones(shape + (9,)), np.zeros(shape + (9,)))
    assert np.allclose(out[m, 4], 1) and np.allclose(out[m, 7], 1 - 1 / 3) and np.allclose(out[m, 8], 1 + 1 / 3)


def test_topology():
    with open(os.path.join(GOLDEN, 'topology.json')) as fh:
        g = json.load(fh)
    for n, v in g['get_xy_size'].items():
        if v is None:
            with pytest.raises(Exception):
                onp.xy_size(int(n))
        else:
            assert tuple(float(a) for a in onp.xy_size(int(n))) == (v[0], v[1])
    for n, lx, ly, cx, cy, nlx, nly in g['local_coords']:
        xs, ys = onp.xy_size(n)
        assert (onp.block_extent(cx, lx, xs), onp.block_extent(cy, ly, ys)) == (nlx, nly)
    for n, lx, ly, cx, cy, gx, gy, xin, yin, lxx, lyy, loc in g['in_process']:
        xs, ys = onp.xy_size(n)
        assert onp.owns(cx, gx, lx, xs) == xin and onp.owns(cy, gy, ly, ys) == yin
        assert onp.to_local(cx, gx, lx, xs) == lxx and onp.to_local(cy, gy, ly, ys) == lyy


@pytest.mark.parametrize('impl', ['c', 'np'])
def test_shear_wave(impl):
    g = load('shear.npz')
    steps = (1, 10, 100, 1000) if impl == 'c' else (1, 10, 100)
    for om in (0.3, 1.0, 1.7):
        rho, u = onp.sinusoidal_velocity_x((100, 50), 0.01)
        if om == 1.0:
            assert np.array_equal(rho, g['rho0']) and np.array_equal(u, g['u0'])
        f = onp.equilibrium(rho, u)
        t = 0
        for tt in steps:
            if impl == 'c':
                f, rho, u = oc.run(f, rho, u, om, oc.periodic(), tt - t)
            else:
                for _ in range(tt - t):
                    f, rho, u = onp.step(f, rho, u, om)
            t = tt
            check_digests(g, f'v_om{om}_t{tt}', f, rho, u)
        if om == 1.0 and impl == 'c':
            assert np.array_equal(f, g['v_f'])
    rho, u = onp.sinusoidal_density_x((50, 50), 0.5, 0.08)
    assert np.array_equal(rho, g['d_rho0'])
    f = onp.equilibrium(rho, u)
    t = 0
    for tt in steps:
        if impl == 'c':
            f, rho, u = oc.run(f, rho, u, 0.8, oc.periodic(), tt - t)
        else:
            for _ in range(tt - t):
                f, rho, u = onp.step(f, rho, u, 0.8)
        t = tt
        check_digests(g, f'd_t{tt}', f, rho, u)


@pytest.mark.parametrize('impl', ['c', 'np'])
def test_couette(impl):
    g = load('couette.npz')
    lx = ly = 100
    rho, u = onp.uniform((lx, ly))
    f = onp.equilibrium(rho, u)
    t = 0
    if impl == 'c':
        sc = oc.couette(0.05, np.mean(rho))
        for tt in (1, 10, 100, 1000, 10000):
            f, rho, u = oc.run(f, rho, u, 1.0, sc, tt - t)
            t = tt
            check_digests(g, f't{tt}', f, rho, u)
            if tt == 100:
                assert np.array_equal(f, g['f100'])
        assert np.array_equal(u[lx // 2, :, 0], g['ux_profile_t10000'])
    else:
        bc = onp.couette_bc(lx, ly, 0.05, np.mean(rho))
        for tt in (1, 10, 100):
            for _ in range(tt - t):
                f, rho, u = onp.step(f, rho, u, 1.0, bc)
            t = tt
            check_digests(g, f't{tt}', f, rho, u)


@pytest.mark.parametrize('impl', ['c', 'np'])
def test_poiseuille(impl):
    g = load('poiseuille.npz')
    lx, ly = 100, 50
    p_in, p_out = float(g['p_in']), float(g['p_out'])
    rho, u = onp.uniform((lx, ly))
    f = onp.equilibrium(rho, u)
    t = 0
    if impl == 'c':
        sc = oc.poiseuille(p_in, p_out)
        for tt in (1, 10, 100, 1000):
            f, rho, u = oc.run(f, rho, u, 1.5, sc, tt - t)
            t = tt
            check_digests(g, f't{tt}', f, rho, u)
        assert np.array_equal(u[1, :, 0], g['ux_profile_x1'])
    else:
        bc = onp.poiseuille_bc(lx, ly, p_in, p_out)
        for tt in (1, 10, 100):
            for _ in range(tt - t):
                f, rho, u = onp.step(f, rho, u, 1.5, bc)
            t = tt
            check_digests(g, f't{tt}', f, rho, u)


KARMAN = dict(lx=420, ly=180, d=40, u0=0.1, rho_in=1.0, nu=0.04)


def test_karman_parallel_path_c():
    """1000 steps of the 1-rank parallel path vs the reference-generated digests, the reference's own 12-sample
    golden (tests/von_karman_vortex_shedding/vel_at_p.npy) and the cluster trace."""
    g = load('karman.npz')
    k = KARMAN
    omega = np.reciprocal(3 * k['nu'] + 0.5)
    rho, u = onp.uniform((k['lx'] + 2, k['ly'] + 2), 1.0, k['u0'], 0.0)
    f = onp.equilibrium(rho, u)
    sc = oc.karman(k['lx'], k['ly'], k['rho_in'], k['u0'], k['d'], ghost=1, probe=(3 * k['lx'] // 4, k['ly'] // 2))
    t = 0
    trace = [u[sc.probe_x, sc.probe_y].copy()]
    for tt in (1, 2, 11, 100, 1000):
        f, rho, u, pr = oc.run(f, rho, u, omega, sc, tt - t, want_probe=True)
        trace.extend(pr)
        t = tt
        check_digests(g, f't{tt}', f, rho, u)
        check_digests(g, f't{tt}_int', f[1:-1, 1:-1], rho[1:-1, 1:-1], u[1:-1, 1:-1])
    trace = np.array(trace)
    assert np.array_equal(trace, g['probe_uxuy'])
    norm = np.array([np.linalg.norm(v) for v in trace])
    assert np.array_equal(norm[:12], load('ref_vel_at_p.npy'))
    assert np.array_equal(norm, load('ref_probe_100.npy')[:1001])
    assert np.array_equal(f[[0, 1, 105, 106, 107, 211, 316, 420, 421]], g['f1000_rows'])


def test_karman_parallel_path_numpy():
    g = load('karman.npz')
    k = KARMAN
    omega = np.reciprocal(3 * k['nu'] + 0.5)
    rho, u = onp.uniform((k['lx'] + 2, k['ly'] + 2), 1.0, k['u0'], 0.0)
    f = onp.equilibrium(rho, u)
    bc = onp.karman_parallel_bc((0, 0), k['lx'], k['ly'], k['lx'], k['ly'], 1, 1, k['rho_in'], k['u0'], k['d'])
    for t in range(1, 12):
        f, rho, u = onp.step(f, rho, u, omega, bc, onp.self_exchange)
        if t in (1, 2, 11):
            check_digests(g, f't{t}', f, rho, u)


def test_karman_serial():
    g = load('karman_serial.npz')
    k = KARMAN
    omega = np.reciprocal(3 * k['nu'] + 0.5)
    rho, u = onp.uniform((k['lx'], k['ly']), 1.0, k['u0'], 0.0)
    f = onp.equilibrium(rho, u)
    sc = oc.karman(k['lx'], k['ly'], k['rho_in'], k['u0'], k['d'], ghost=0)
    t = 0
    for tt in (1, 11, 200):
        f, rho, u = oc.run(f, rho, u, omega, sc, tt - t)
        t = tt
        check_digests(g, f't{tt}', f, rho, u)
    # numpy flavour, 11 steps
    rho, u = onp.uniform((k['lx'], k['ly']), 1.0, k['u0'], 0.0)
    f = onp.equilibrium(rho, u)
    bc = onp.karman_serial_bc(k['lx'], k['ly'], k['rho_in'], k['u0'], k['d'])
    for t in range(1, 12):
        f, rho, u = onp.step(f, rho, u, omega, bc)
    check_digests(g, 't11', f, rho, u)


@pytest.mark.parametrize('size', [2, 4, 6, 14])
def test_karman_ranks_numpy(size):
    """k ranks emulated in-process by the oracle == the reference run on k (thread-)ranks, ghost rings included."""
    g = load('karman_ranks.npz')
    k = KARMAN
    lx, ly = k['lx'], k['ly']
    omega = np.reciprocal(3 * k['nu'] + 0.5)
    xs, ys = onp.xy_size(size)
    F, R, U, bcs = {}, {}, {}, {}
    for cx in range(xs):
        for cy in range(ys):
            nlx, nly = onp.block_extent(cx, lx, xs), onp.block_extent(cy, ly, ys)
            R[(cx, cy)], U[(cx, cy)] = onp.uniform((nlx + 2, nly + 2), 1.0, k['u0'], 0.0)
            F[(cx, cy)] = onp.equilibrium(R[(cx, cy)], U[(cx, cy)])
            bcs[(cx, cy)] = onp.karman_parallel_bc((cx, cy), nlx, nly, lx, ly, xs, ys, k['rho_in'], k['u0'], k['d'])
    for _ in range(11):
        F, R, U = onp.step_blocks(F, R, U, omega, bcs, xs, ys)
    G = np.zeros((lx, ly, 9))
    for (cx, cy), fb in F.items():
        r = cx * ys + cy
        assert sha(fb) == str(g[f'n{size}_r{r}_f_full'])
        assert sha(R[(cx, cy)]) == str(g[f'n{size}_r{r}_rho_full'])
        assert sha(U[(cx, cy)]) == str(g[f'n{size}_r{r}_u_full'])
        x0, y0 = onp.block_origin(cx, lx, xs), onp.block_origin(cy, ly, ys)
        G[x0:x0 + fb.shape[0] - 2, y0:y0 + fb.shape[1] - 2] = fb[1:-1, 1:-1]
    assert sha(G) == str(g['serial_f11']) == str(g[f'n{size}_f11'])
