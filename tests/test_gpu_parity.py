"""Parity of the CUDA path (through the C-ABI, via the reference-shaped Python API) against the pinned oracle and
the reference-generated golden fixtures. Bit-exact everywhere (every operation is an IEEE add/mul/div/sqrt in the
reference's order); the bar BASELINE.json states is 1e-12 relative after 1000 steps (tests/helpers.py:RTOL).
Runs on the B200 box: `pytest -m gpu`."""
import os

import numpy as np
import pytest

from tests.helpers import assert_parity, sha

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope='module')
def P():
    import lattice_boltzmann_parallel_solver_b200 as pkg
    from lattice_boltzmann_parallel_solver_b200 import _native
    _native.device()   # raises loudly when there is no GPU / no library
    yield pkg
    pkg.lattice_boltzmann_method.release_lattices()


@pytest.fixture(scope='module')
def oracle():
    from oracle import lbm_c, lbm_numpy

    class O:
        c = lbm_c
        np = lbm_numpy
    return O


def check_digests(g, tag, f, rho, u):
    assert sha(np.asarray(f)) == str(g[tag + '_f']), tag + ' f'
    assert sha(np.asarray(rho)) == str(g[tag + '_rho']), tag + ' rho'
    assert sha(np.asarray(u)) == str(g[tag + '_u']), tag + ' u'


# ---------------------------------------------------------------------------------------------------------
# a2-a5: stateless kernels
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', ['a', 'b'])
def test_kernels_vs_reference_golden(P, n):
    L = P.lattice_boltzmann_method
    g = load('kernels.npz')
    rho, u, f = g[n + '_rho'], g[n + '_u'], g[n + '_f']
    assert_parity(L.equilibrium_distr_func(rho, u), g[n + '_feq'], 'feq')
    assert_parity(L.compute_density(f), g[n + '_density'], 'density')
    assert_parity(L.compute_velocity_field(rho, f), g[n + '_velocity'], 'velocity')
    assert_parity(L.streaming(f), g[n + '_stream'], 'streaming')
    f2, r2, u2 = L.lattice_boltzmann_step(f, rho, u, 1.3)
    assert_parity(f2, g[n + '_step_f'], 'step f')
    assert_parity(r2, g[n + '_step_rho'], 'step rho')
    assert_parity(u2, g[n + '_step_u'], 'step u')


def test_hand_expanded_division_and_sqrt_equal_ieee(P):
    """lbm_device.cuh expands u = j / rho by hand (one reciprocal for both quotients) and special-cases exact zeros:
    every quotient and root must carry the bits of the compiler's IEEE __ddiv_rn / __dsqrt_rn, over random bit
    patterns, the physical range, the range-test thresholds, signed zeros and exact quotients."""
    import ctypes as C
    import struct
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    lib = N.load()
    out = (C.c_uint64 * 7)()
    for seed in (1, 20261017, 0xdeadbeefcafe):
        N.check(lib.lbm_selftest_arith(N.device(), 1 << 27, seed, out))
        as_f = [struct.unpack('<d', struct.pack('<Q', v))[0] for v in out[2:5]]
        assert out[0] == 0, f'{out[0]} quotients differ from __ddiv_rn, first: {as_f[0]!r} / {as_f[1]!r} ({out[2]:#x}, {out[3]:#x})'
        assert out[1] == 0, f'{out[1]} roots differ from __dsqrt_rn, first: sqrt({as_f[2]!r}) ({out[4]:#x})'
        # the branch-free variants must answer the bulk themselves (physical range, exact quotients, zeros, ...)
        assert out[5] > (1 << 27) // 3 and out[6] > (1 << 27) // 4, (out[5], out[6])


def test_equilibrium_1d_inputs(P, oracle):
    """boundary_conditions.py:338 calls it with (ly,), (ly,2) and gets (1, ly, 9)."""
    L = P.lattice_boltzmann_method
    rng = np.random.default_rng(5)
    rho, u = rng.uniform(0.9, 1.1, 13), rng.uniform(-0.1, 0.1, (13, 2))
    out = L.equilibrium_distr_func(rho, u)
    assert out.shape == (1, 13, 9)
    assert_parity(out[0], oracle.np.equilibrium(rho, u), '1-D feq')


def test_step_asserts(P):
    L = P.lattice_boltzmann_method
    f, rho, u = np.ones((4, 4, 9)), np.ones((4, 4)), np.zeros((4, 4, 2))
    for om in (0, 2, -1, 2.5):
        with pytest.raises(AssertionError):
            L.lattice_boltzmann_step(f, rho, u, om)
    with pytest.raises(AssertionError):
        L.lattice_boltzmann_step(f, np.ones((4, 5)), u, 1.0)
    with pytest.raises(AssertionError):
        L.compute_density(np.ones((4, 4, 8)))
    with pytest.raises(TypeError):
        L.lattice_boltzmann_step(f, rho, u, 1.0, boundary=lambda *a: a[1])


# ---------------------------------------------------------------------------------------------------------
# reference unit tests re-pointed at the new module (tests/test_*.py of the reference)
# ---------------------------------------------------------------------------------------------------------
def test_reference_unit_invariants(P):
    L = P.lattice_boltzmann_method
    # test_density_computation.py:12-18 / test_velocity_computation.py:12-20
    f = np.ones((10, 10, 9)) / 9
    assert round(float(np.sum(L.compute_density(f))), 1) == 100.0
    assert round(float(np.sum(L.compute_velocity_field(L.compute_density(f), f))), 1) == 0.0
    # test_streaming_func.py:12-38
    rng = np.random.default_rng(0)
    for g in (np.ones((12, 9, 9)), rng.uniform(0, 1, (12, 9, 9))):
        assert abs(float(np.sum(L.streaming(g)) - np.sum(g))) < 1e-9
    one = np.zeros((12, 9, 9))
    one[3, 4, 5] = 1.0
    assert L.streaming(one)[4, 5, 5] == 1.0
    # test_navier_stokes_eq.py:12-51
    rho = rng.uniform(0.5, 1.5, (10, 10))
    u = rng.uniform(-0.1, 0.1, (10, 10, 2))
    feq = L.equilibrium_distr_func(rho, u)
    assert np.allclose(feq.sum(-1), rho)
    assert np.allclose(feq @ L.get_velocity_sets(), rho[..., None] * u)
    assert L.get_velocity_sets().sum() == 0


def test_mass_preservation_10000_steps(P):
    """tests/test_mass_preservation.py:22-34: 50x50, omega 0.5, point perturbation, 10 000 periodic steps."""
    L = P.lattice_boltzmann_method
    rho = np.ones((50, 50)) * 0.5
    rho[25, 25] = 0.6
    u = np.zeros((50, 50, 2))
    f = L.equilibrium_distr_func(rho, u)
    m0 = float(np.sum(rho))
    for _ in range(10000):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 0.5)
    assert round(float(np.sum(np.asarray(rho))), 1) == round(m0, 1)
    assert abs(float(np.sum(np.asarray(f))) - m0) < 1e-8


# ---------------------------------------------------------------------------------------------------------
# a6-a11: boundary closures called directly (tests/test_boundary_conditions.py of the reference)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', ['s', 'r'])
def test_bc_primitives_vs_reference_golden(P, n):
    B, BU = P.boundary_conditions, P.boundary_utils
    g = load('bc_primitives.npz')
    f_pre, f_post, f_prev, rho, u = (g[n + k] for k in ('_f_pre', '_f_post', '_f_prev', '_rho', '_u'))
    nx, ny = rho.shape

    def edge(w):
        m = np.zeros((nx, ny), dtype=bool)
        if w == 'x0':
            m[0, :] = True
        elif w == 'x1':
            m[-1, :] = True
        elif w == 'y0':
            m[:, 0] = True
        else:
            m[:, -1] = True
        return m

    for w in ('x0', 'x1', 'y0', 'y1'):
        assert_parity(B.rigid_wall(edge(w))(f_pre.copy(), f_post.copy()), g[f'{n}_rigid_{w}'], 'rigid ' + w)
        assert_parity(B.moving_wall(edge(w), np.array([0.05, -0.02]), 1.03)(f_pre.copy(), f_post.copy()),
                      g[f'{n}_moving_{w}'], 'moving ' + w)
    assert_parity(B.inlet((nx, ny), 1.02, 0.1)(f_post.copy()), g[n + '_inlet'], 'inlet')
    assert_parity(B.outlet()(f_prev.copy(), f_post.copy()), g[n + '_outlet'], 'outlet')
    m = edge('x0') | edge('x1')
    assert_parity(B.periodic_with_pressure_variations(m, 0.3345, 0.3321)(f_pre.copy(), rho, u), g[n + '_pbc_x'], 'pbc')
    pm = g[n + '_plate_mask'].copy()
    assert_parity(B.rigid_object(pm)(f_pre.copy(), f_post.copy()), g[n + '_plate'], 'plate')
    assert np.array_equal(pm, g[n + '_plate_mask_after'])   # the reference mutates the caller's mask
    # in-place semantics: the closure returns the array it was given
    fp = f_post.copy()
    assert B.rigid_wall(edge('y1'))(f_pre, fp) is fp
    # bundles called directly
    assert_parity(BU.couette_flow_boundary_conditions(nx, ny, 0.05, 1.0)(f_pre.copy(), f_post.copy(), rho, u, f_prev),
                  g[n + '_couette'], 'couette bundle')
    fpre = f_pre.copy()
    assert_parity(BU.poiseuille_flow_boundary_conditions(nx, ny, 0.3345, 0.3321)(fpre, f_post.copy(), rho, u, f_prev),
                  g[n + '_poiseuille'], 'poiseuille bundle')
    assert_parity(fpre, g[n + '_poiseuille_fpre_after'], 'poiseuille f_pre after')


def test_reference_bc_known_answers(P):
    """The literal constants of tests/test_boundary_conditions.py:62-121."""
    B = P.boundary_conditions
    shape = (10, 10)
    m = np.zeros(shape, dtype=bool)
    m[:, -1] = True
    out = B.moving_wall(m, np.array([2, 0]), 1)(np.ones(shape + (9,)), np.zeros(shape + (9,)))
    assert np.allclose(out[m, 4], 1) and np.allclose(out[m, 7], 1 - 1 / 3) and np.allclose(out[m, 8], 1 + 1 / 3)
    for i in (0, 1, 2, 3, 5, 6):
        assert np.allclose(out[m, i], 0)
    u = np.zeros(shape + (2,))
    u[..., 0] = 0.1
    b = np.zeros(shape, dtype=bool)
    b[0, :] = True
    b[-1, :] = True
    out = B.periodic_with_pressure_variations(b, 1, 0.1)(np.ones(shape + (9,)), np.ones(shape), u)
    assert np.allclose(out[0, :, 1], 1.29555) and np.allclose(out[0, :, 5], 1.073888) and np.allclose(out[0, :, 8], 1.073888)
    assert np.allclose(out[-1, :, 3], 0.943222) and np.allclose(out[-1, :, 6], 0.9858) and np.allclose(out[-1, :, 7], 0.9858)
    assert np.allclose(out[1:-1], 1)


# ---------------------------------------------------------------------------------------------------------
# configs 1-4 (BASELINE.json) through the reference's driver loops
# ---------------------------------------------------------------------------------------------------------
def test_config1_shear_wave_omega_sweep(P, oracle):
    L = P.lattice_boltzmann_method
    g = load('shear.npz')
    for om in (0.3, 1.0, 1.7):
        rho, u = oracle.np.sinusoidal_velocity_x((100, 50), 0.01)
        f = L.equilibrium_distr_func(rho, u)
        amp = []
        for t in range(1, 1001):
            f, rho, u = L.lattice_boltzmann_step(f, rho, u, om)
            if t <= 200:   # the driver's per-step whole-field reduction (experiments.py:188-193), on the device
                vmin, vmax = np.amin(u), np.amax(u)
                amp.append(np.abs(vmin) if np.abs(vmin) > np.abs(vmax) else np.abs(vmax))
            if t in (1, 10, 100, 1000):
                check_digests(g, f'v_om{om}_t{t}', f, rho, u)
        assert np.array_equal(np.array(amp), g[f'v_om{om}_amp'][:200])
    rho, u = oracle.np.sinusoidal_density_x((50, 50), 0.5, 0.08)
    f = L.equilibrium_distr_func(rho, u)
    for t in range(1, 1001):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 0.8)
        if t in (1, 10, 100, 1000):
            check_digests(g, f'd_t{t}', f, rho, u)


def test_omega_change_between_calls(P, oracle):
    """experiments.py:171-180 sweeps omega; the speculative collision of the resident state must be redone."""
    L = P.lattice_boltzmann_method
    rho, u = oracle.np.sinusoidal_velocity_x((32, 48), 0.05)
    f = L.equilibrium_distr_func(rho, u)
    fo, ro, uo = f.copy(), rho.copy(), u.copy()
    omegas = [0.7, 0.7, 1.4, 1.4, 0.2, 1.9, 1.9]
    for om in omegas:
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, om)
        fo, ro, uo = oracle.c.run(fo, ro, uo, om, oracle.c.periodic(), 1)
    assert_parity(f, fo, 'f')
    assert_parity(rho, ro, 'rho')
    assert_parity(u, uo, 'u')


@pytest.mark.parametrize('mode', ['mask', 'edge'])
def test_config2_couette(P, oracle, mode):
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    g = load('couette.npz')
    lx = ly = 100
    rho, u = oracle.np.uniform((lx, ly))
    f = P.lattice_boltzmann_method.equilibrium_distr_func(rho, u)
    bc = P.boundary_utils.couette_flow_boundary_conditions(lx, ly, 0.05, np.mean(rho))
    lat = Lattice(lx, ly, bc.kind_map((lx, ly)), bc_mode=N.BC_MASK if mode == 'mask' else N.BC_EDGE)
    lat.load(f, rho, u, 1.0)
    t = 0
    for tt in (1, 10, 100, 1000, 10000):
        lat.run(tt - t)
        t = tt
        fo, ro, uo = lat.fields()
        check_digests(g, f't{tt}', fo, ro, uo)
    assert np.array_equal(uo[lx // 2, :, 0], g['ux_profile_t10000'])
    lat.close()


def test_config2_couette_python_loop(P, oracle):
    g = load('couette.npz')
    L = P.lattice_boltzmann_method
    lx = ly = 100
    rho, u = oracle.np.uniform((lx, ly))
    f = L.equilibrium_distr_func(rho, u)
    bc = P.boundary_utils.couette_flow_boundary_conditions(lx, ly, 0.05, np.mean(rho))
    kept = []
    for t in range(1, 101):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.0, bc)
        if t in (1, 10):
            kept.append((t, f, rho, u))     # experiments.py:254 keeps every step's velocity: handles must stay valid
    check_digests(g, 't100', f, rho, u)
    assert_parity(f, g['f100'], 'f100')
    for t, a, b, c in kept:
        check_digests(g, f't{t}', a, b, c)


@pytest.mark.parametrize('mode', ['mask', 'edge'])
def test_config3_poiseuille(P, oracle, mode):
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    g = load('poiseuille.npz')
    lx, ly = 100, 50
    rho, u = oracle.np.uniform((lx, ly))
    f = P.lattice_boltzmann_method.equilibrium_distr_func(rho, u)
    bc = P.boundary_utils.poiseuille_flow_boundary_conditions(lx, ly, float(g['p_in']), float(g['p_out']))
    lat = Lattice(lx, ly, bc.kind_map((lx, ly)), bc_mode=N.BC_MASK if mode == 'mask' else N.BC_EDGE)
    lat.load(f, rho, u, 1.5)
    t = 0
    for tt in (1, 10, 100, 1000):
        lat.run(tt - t)
        t = tt
        fo, ro, uo = lat.fields()
        check_digests(g, f't{tt}', fo, ro, uo)
        if tt == 100:
            assert_parity(fo, g['f100'], 'f100')
    assert np.array_equal(uo[1, :, 0], g['ux_profile_x1'])
    assert np.array_equal(ro[:, ly // 2], g['rho_centerline'])
    lat.close()


KARMAN = dict(lx=420, ly=180, d=40, u0=0.1, rho_in=1.0, nu=0.04)


def test_config4_karman_parallel_path_1000_steps(P, oracle):
    """The loop of experiments.py:650-704 / tests/test_parallelization_von_karman.py:18-57 on one rank: ghost ring,
    communication(), parallel BC bundle, per-step probe read. 1000 steps; fields, ghost ring included, and the
    probe trace must equal the reference's bit for bit (incl. its own 12-sample golden and the cluster trace)."""
    L, BU, PU = P.lattice_boltzmann_method, P.boundary_utils, P.parallelization_utils
    from lattice_boltzmann_parallel_solver_b200 import dist
    g = load('karman.npz')
    k = KARMAN
    lx, ly = k['lx'], k['ly']
    omega = np.reciprocal(3 * k['nu'] + 0.5)
    comm = dist.WorldComm()
    x_size, y_size = PU.get_xy_size(comm.Get_size())
    cart = comm.Create_cart(dims=[x_size, y_size], periods=[True, True], reorder=False)
    c = cart.Get_coords(comm.Get_rank())
    nlx, nly = PU.get_local_coords(c, lx, ly, x_size, y_size)
    rho, u = oracle.np.uniform((nlx + 2, nly + 2), 1.0, k['u0'], 0.0)
    f = L.equilibrium_distr_func(rho, u)
    pc, px, py = PU.global_coord_to_local_coord(c, 3 * lx // 4, ly // 2, lx, ly, x_size, y_size)
    vel_at_p = [np.linalg.norm(u[px, py, ...])]
    comps = [u[px, py].copy()]
    bc = BU.parallel_von_karman_boundary_conditions(c, nlx, nly, lx, ly, x_size, y_size, k['rho_in'], k['u0'], k['d'])
    com = PU.communication(cart)
    for t in range(1, 1001):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, omega, bc, com)
        v = u[px, py, ...]
        comps.append(np.array(v))
        vel_at_p.append(np.linalg.norm(v))
        if t in (1, 2, 11, 100, 1000):
            check_digests(g, f't{t}', f, rho, u)
            F, R, U = np.asarray(f), np.asarray(rho), np.asarray(u)
            check_digests(g, f't{t}_int', F[1:-1, 1:-1], R[1:-1, 1:-1], U[1:-1, 1:-1])
    assert np.array_equal(np.array(comps), g['probe_uxuy'])
    assert np.array_equal(np.array(vel_at_p)[:12], load('ref_vel_at_p.npy'))
    assert np.array_equal(np.array(vel_at_p), load('ref_probe_100.npy')[:1001])


def test_config4_karman_native_run_with_probe(P, oracle):
    """Same scenario through the native multi-step API: 1000 steps in one call, probe ring read afterwards."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    g = load('karman.npz')
    k = KARMAN
    lx, ly = k['lx'], k['ly']
    omega = float(np.reciprocal(3 * k['nu'] + 0.5))
    rho, u = oracle.np.uniform((lx + 2, ly + 2), 1.0, k['u0'], 0.0)
    f = P.lattice_boltzmann_method.equilibrium_distr_func(rho, u)
    bc = P.boundary_utils.parallel_von_karman_boundary_conditions([0, 0], lx, ly, lx, ly, 1, 1, k['rho_in'], k['u0'], k['d'])
    lat = Lattice(lx + 2, ly + 2, bc.kind_map((lx + 2, ly + 2)), ghost=(1, 1))
    lat.connect_self_periodic()
    lat.probe(3 * lx // 4 + 1, ly // 2 + 1, capacity=2048)
    lat.load(f, rho, u, omega)
    lat.run(1000)
    trace = lat.probe_read(1, 1000)
    assert np.array_equal(trace, g['probe_uxuy'][1:])
    fo, ro, uo = lat.fields()
    check_digests(g, 't1000', fo, ro, uo)
    mn_r, mx_r, mn_u, mx_u = lat.minmax()
    assert (mn_r, mx_r, mn_u, mx_u) == (ro.min(), ro.max(), uo.min(), uo.max())
    lat.close()


@pytest.mark.parametrize('calls', [(64,), (96, 32), (33, 31), (1, 1, 62), (40, 64, 1)])
def test_ghost_ring_results_after_calls_that_end_on_a_graph_replay(P, oracle, calls):
    """The ghost-ring snapshot that materialisation reads is taken on the LAST step of a call only — also when that step
    sits at the end of a replayed CUDA graph (second graph variant) — and the probe's time word lags one launch behind:
    fields and probe samples after every call equal the C oracle (parallel von Karman block with its own ghost ring)."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    lx, ly, d = 62, 40, 8
    rng = np.random.default_rng(21)
    shape = (lx + 2, ly + 2)
    rho = rng.uniform(0.9, 1.1, shape); u = rng.uniform(-0.05, 0.05, shape + (2,)); f = oracle.np.equilibrium(rho, u)
    bc = P.boundary_utils.parallel_von_karman_boundary_conditions([0, 0], lx, ly, lx, ly, 1, 1, 1.0, 0.1, d)
    lat = Lattice(*shape, bc.kind_map(shape), ghost=(1, 1))
    lat.connect_self_periodic()
    px, py = 3 * lx // 4 + 1, ly // 2 + 1
    lat.probe(px, py, capacity=512)
    lat.load(f, rho, u, 1.6)
    scen = oracle.c.karman(lx, ly, 1.0, 0.1, d, ghost=1)
    state, t = (f, rho, u), 0
    samples = []
    for n in calls:
        lat.run(n)
        for _ in range(n):
            state = oracle.c.run(*state, 1.6, scen, 1)
            samples.append(state[2][px, py].copy())
        t += n
        assert np.array_equal(lat.probe_read(t - n + 1, n), np.array(samples[t - n:t])), (calls, t)
        got = lat.fields()
        for a, b, name in zip(got, state, ('f', 'rho', 'u')):
            assert np.array_equal(a, b), (calls, t, name)
    lat.close()


def test_karman_serial_rigid_object(P, oracle):
    """milestoneQuickFunctionCalls.py:304-320 — inlet, outlet and rigid_object composed by hand, no ghost ring."""
    L, B, BU = P.lattice_boltzmann_method, P.boundary_conditions, P.boundary_utils
    g = load('karman_serial.npz')
    k = KARMAN
    lx, ly, d = k['lx'], k['ly'], k['d']
    omega = np.reciprocal(3 * k['nu'] + 0.5)
    plate = np.zeros((lx, ly))
    plate[lx // 4, ly // 2 - d // 2:ly // 2 + d // 2] = 1
    bundle = BU.BoundaryBundle('milestone_6', (lx, ly))
    bundle.add(B.inlet((lx, ly), k['rho_in'], k['u0'])).add(B.outlet()).add(B.rigid_object(plate.astype(bool)))
    rho, u = oracle.np.uniform((lx, ly), 1.0, k['u0'], 0.0)
    f = L.equilibrium_distr_func(rho, u)
    for t in range(1, 201):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, omega, bundle)
        if t in (1, 11, 200):
            check_digests(g, f't{t}', f, rho, u)
    assert_parity(np.asarray(u)[[0, 1, 105, 106, 315, 418, 419]], g['u200_rows'], 'u rows')


# ---------------------------------------------------------------------------------------------------------
# fuzz against the C oracle on ragged sizes and random fields
# ---------------------------------------------------------------------------------------------------------
def random_state(oracle, shape, seed):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.9, 1.1, shape)
    ang = rng.uniform(0, 2 * np.pi, shape)
    mag = rng.uniform(0, 0.05, shape)
    u = np.dstack([mag * np.cos(ang), mag * np.sin(ang)])
    f = oracle.np.equilibrium(rho, u) * rng.uniform(0.98, 1.02, shape + (9,))   # off equilibrium
    return f, rho, u


@pytest.mark.parametrize('shape', [(1, 1), (1, 7), (5, 1), (2, 2), (3, 17), (37, 23), (130, 257), (64, 1024), (300, 1000)])
def test_fuzz_periodic(P, oracle, shape):
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    f, rho, u = random_state(oracle, shape, 11)
    lat = Lattice(*shape)
    lat.load(f, rho, u, 1.23)
    lat.run(20)
    got = lat.fields()
    ref = oracle.c.run(f, rho, u, 1.23, oracle.c.periodic(), 20)
    for a, b, n in zip(got, ref, 'f rho u'.split()):
        assert_parity(a, b, f'{shape} {n}')
    # sub-rectangle read
    x0, x1 = 0, max(1, shape[0] // 2)
    y0, y1 = shape[1] // 3, shape[1]
    part = lat.fields(region=(x0, x1, y0, y1))
    for a, b in zip(part, ref):
        assert_parity(a, b[x0:x1, y0:y1], f'{shape} region')
    lat.close()


@pytest.mark.parametrize('shape,mode', [((24, 31), 'mask'), ((24, 31), 'edge'), ((130, 70), 'mask'), ((130, 70), 'edge')])
def test_fuzz_scenarios(P, oracle, shape, mode):
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    BU = P.boundary_utils
    nx, ny = shape
    bm = N.BC_MASK if mode == 'mask' else N.BC_EDGE
    f, rho, u = random_state(oracle, shape, 3)
    for name in ('couette', 'poiseuille'):
        if name == 'couette':
            bc = BU.couette_flow_boundary_conditions(nx, ny, 0.03, 1.01)
            sc = oracle.c.couette(0.03, 1.01)
        else:
            bc = BU.poiseuille_flow_boundary_conditions(nx, ny, 0.3345, 0.3321)
            sc = oracle.c.poiseuille(0.3345, 0.3321)
        lat = Lattice(nx, ny, bc.kind_map(shape), bc_mode=bm)
        lat.load(f, rho, u, 0.9)
        lat.run(30)
        ref = oracle.c.run(f, rho, u, 0.9, sc, 30)
        for a, b, n in zip(lat.fields(), ref, 'f rho u'.split()):
            assert_parity(a, b, f'{name} {shape} {mode} {n}')
        lat.close()
    # von Karman with ghost ring
    lx, ly = nx - 2, ny - 2
    bc = BU.parallel_von_karman_boundary_conditions([0, 0], lx, ly, lx, ly, 1, 1, 1.0, 0.1, 8)
    lat = Lattice(nx, ny, bc.kind_map(shape), ghost=(1, 1), bc_mode=bm)
    lat.connect_self_periodic()
    lat.load(f, rho, u, 1.6)
    lat.run(30)
    ref = oracle.c.run(f, rho, u, 1.6, oracle.c.karman(lx, ly, 1.0, 0.1, 8, ghost=1), 30)
    for a, b, n in zip(lat.fields(), ref, 'f rho u'.split()):
        assert_parity(a, b, f'karman {shape} {mode} {n}')
    lat.close()


def test_device_init_matches_upload(P, oracle):
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    shape = (48, 80)
    rho, u = oracle.np.sinusoidal_velocity_x(shape, 0.01)
    f = oracle.np.equilibrium(rho, u)
    a = Lattice(*shape)
    a.load(f, rho, u, 1.1)
    b = Lattice(*shape)
    b.load_equilibrium(1.1, ux_y=u[0, :, 0])
    a.run(7)
    b.run(7)
    for x, y in zip(a.fields(), b.fields()):
        assert_parity(x, y, 'device init')
    rho, u = oracle.np.sinusoidal_density_x(shape, 0.5, 0.08)
    b.load_equilibrium(0.6, rho_x=rho[:, 0])
    b.run(5)
    ref = oracle.c.run(oracle.np.equilibrium(rho, u), rho, u, 0.6, oracle.c.periodic(), 5)
    for x, y in zip(b.fields(), ref):
        assert_parity(x, y, 'device init density')
    a.close()
    b.close()


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json's full size through size-independent properties
# ---------------------------------------------------------------------------------------------------------
def test_full_size_16384_rows_equal_thin_lattice(P, oracle):
    """The throughput workload (16384x16384 periodic shear wave) is invariant along x, so every
    row must equal — bit for bit — the rows of a 4x16384 lattice the C oracle can run. Also: total mass is
    conserved and the fields are exactly x-invariant."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    n, steps = 16384, 24
    y = np.arange(n)
    prof = 0.01 * np.sin(np.divide(2 * np.pi * y, n))
    lat = Lattice(n, n)
    lat.load_equilibrium(1.0, ux_y=prof)
    lat.run(steps)
    rho, u = oracle.np.sinusoidal_velocity_x((4, n), 0.01)
    assert np.array_equal(u[0, :, 0], prof)
    ref = oracle.c.run(oracle.np.equilibrium(rho, u), rho, u, 1.0, oracle.c.periodic(), steps)
    for x0 in (0, 1, 8191, 16380):
        got = lat.fields(region=(x0, x0 + 4, 0, n))
        for a, b, nm in zip(got, ref, 'f rho u'.split()):
            assert_parity(a, b, f'rows {x0}.. {nm}')
    mn_r, mx_r, mn_u, mx_u = lat.minmax()
    assert mn_r == ref[1].min() and mx_r == ref[1].max() and mn_u == ref[2].min() and mx_u == ref[2].max()
    lat.close()


def test_full_size_16384_x_periodic_field_default_geometry(P):
    """The headline launch geometry (default depth and segment length at 16384^2) on a field that varies along BOTH
    axes: rho(x) with period 64 on top of the shear wave, so that the solution equals a 64 x 16384 lattice the C
    oracle runs in a second (the same job bench.py runs at every N and reports as `parity`). A row mix-up inside a
    segment, a wrong ring slot or a wrong column shift cannot hide here."""
    import bench
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    n = 16384
    prof = 0.01 * np.sin(np.divide(2 * np.pi * np.arange(n), n))
    lat = Lattice(n, n)
    for steps in (12, 13):            # ends on a three-step pass / on a one-step launch
        res = bench.run_parity(lat, 1, 0, n, n, prof, steps, lambda: None, lambda v: v, lambda v: v)
        assert res['mismatches'] == 0 and res['rows'] == 12, res
    lat.close()


# ---------------------------------------------------------------------------------------------------------
# two steps per pass (temporal blocking) == two one-step passes
# ---------------------------------------------------------------------------------------------------------
def _lattice_with_env(shape, env, **kw):
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return Lattice(*shape, **kw)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize('shape', [(1024, 1024), (4100, 258), (513, 2050)])
def test_two_steps_per_pass_equals_single_steps(P, oracle, shape):
    f, rho, u = random_state(oracle, shape, 21)
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    fused = Lattice(*shape)
    fused.set_option('fused_depth', 2)
    plain = _lattice_with_env(shape, {'LBM_NO_FUSED': '1'}) if shape[0] == 1024 else Lattice(*shape)
    plain.set_option('fused', 0)
    px, py = shape[0] // 3, shape[1] - 2
    for lat in (fused, plain):
        lat.probe(px, py, capacity=64)
        lat.load(f, rho, u, 1.37)
    l0 = fused.launches
    for n in (7, 1, 8, 2):            # odd and even counts: 3 pairs + 1, a lone step, 4 pairs, 1 pair (a fluid lattice
        fused.run(n)                  # may end a call on a pass: results are then rebuilt by re-running it in FINAL mode)
        plain.run(n)
    assert fused.launches - l0 == (3 + 1) + 1 + 4 + 1, 'the two-step kernel was not used'
    for a, b, nm in zip(fused.fields(), plain.fields(), 'f rho u'.split()):
        assert_parity(a, b, f'{shape} {nm}')
    assert_parity(fused.probe_read(1, 18), plain.probe_read(1, 18), 'probe ring')
    if shape == (1024, 1024):
        ref = oracle.c.run(f, rho, u, 1.37, oracle.c.periodic(), 18)
        for a, b, nm in zip(fused.fields(), ref, 'f rho u'.split()):
            assert_parity(a, b, f'vs oracle {nm}')
    fused.close()
    plain.close()


# ---------------------------------------------------------------------------------------------------------
# D steps per pass (k_stepNx, D = 2, 3, 4) == D one-step passes == the C oracle, on random (x- and y-varying) fields
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('depth', [2, 3, 4])
@pytest.mark.parametrize('shape', [(1024, 1024), (4100, 258), (513, 2050)])
def test_deep_passes_equal_single_steps(P, oracle, shape, depth):
    f, rho, u = random_state(oracle, shape, 31 + depth)
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    fused, plain = Lattice(*shape), Lattice(*shape)
    fused.set_option('fused_depth', depth)
    fused.set_option('deep2', 1)
    plain.set_option('fused', 0)
    px, py = shape[0] // 3, shape[1] - 2
    for lat in (fused, plain):
        lat.probe(px, py, capacity=64)
        lat.load(f, rho, u, 1.37)
    total = 0
    for n in (2 * depth + 1, 1, 3 * depth, 2, depth + 2):
        fused.run(n)
        plain.run(n)
        total += n
    assert fused.launches < plain.launches, 'the multi-step kernel was not used'
    for a, b, nm in zip(fused.fields(), plain.fields(), 'f rho u'.split()):
        assert_parity(a, b, f'{shape} depth {depth} {nm}')
    assert_parity(fused.probe_read(1, total), plain.probe_read(1, total), 'probe ring')
    if shape == (1024, 1024):
        ref = oracle.c.run(f, rho, u, 1.37, oracle.c.periodic(), total)
        for a, b, nm in zip(fused.fields(), ref, 'f rho u'.split()):
            assert_parity(a, b, f'vs oracle {nm}')
    fused.close()
    plain.close()


@pytest.mark.parametrize('depth,seg,tail', [(2, 16, 1), (2, 64, 0), (3, 16, 0), (3, 32, 1), (3, 64, 1), (3, 128, 0), (3, 256, 1),
                                            (4, 64, 0)])
def test_deep_passes_segment_lengths(P, oracle, depth, seg, tail):
    """The launch geometry of the headline (long segments, which pick_seg only chooses from 8192^2 up) on a random
    field at 4096 x 512, against the C oracle: a row mix-up inside a segment cannot hide here."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    shape = (4096, 512)
    f, rho, u = random_state(oracle, shape, 77)
    lat = Lattice(*shape)
    lat.set_option('fused_depth', depth)
    lat.set_option('deep2', 1)
    lat.set_option('fused_seg', seg)
    lat.load(f, rho, u, 0.9)
    steps = 2 * depth + tail          # tail = 0: the call ENDS on a pass and fields() re-runs it in FINAL mode
    l0 = lat.launches
    lat.run(steps)
    assert lat.launches - l0 == 2 + tail, 'two multi-step passes (and one single step) expected'
    ref = oracle.c.run(f, rho, u, 0.9, oracle.c.periodic(), steps)
    for a, b, nm in zip(lat.fields(), ref, 'f rho u'.split()):
        assert_parity(a, b, f'depth {depth} seg {seg} {nm}')
    got = lat.fields(region=(1000, 1003, 250, 259))          # a sub-rectangle that straddles two column strips
    for a, b, nm in zip(got, ref, 'f rho u'.split()):
        assert_parity(a, b[1000:1003, 250:259], f'region {nm}')
    mn_r, mx_r, mn_u, mx_u = lat.minmax()
    assert (mn_r, mx_r, mn_u, mx_u) == (ref[1].min(), ref[1].max(), ref[2].min(), ref[2].max())
    lat.close()


def test_omega_change_after_a_multi_step_pass(P, oracle):
    """experiments.py:171-180 sweeps omega over one resident state: the collision that ended the previous call is
    redone with the new omega — after a multi-step pass by re-running it with a different omega for its last level."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    shape = (2048, 512)
    f, rho, u = random_state(oracle, shape, 91)
    lat = Lattice(*shape)
    lat.load(f, rho, u, 0.7)
    lat.run(6, 0.7)                   # two three-step passes; the call ends on a pass
    lat.run(5, 1.6)                   # redo of the last pass (levels 1-2 with 0.7, level 3 with 1.6), then 3 + 2
    ref = oracle.c.run(f, rho, u, 0.7, oracle.c.periodic(), 6)
    ref = oracle.c.run(*ref, 1.6, oracle.c.periodic(), 5)
    for a, b, nm in zip(lat.fields(), ref, 'f rho u'.split()):
        assert_parity(a, b, f'omega change {nm}')
    lat.run(4, 1.1)                   # the previous call ended on a TWO-step pass (k_step2x): redone by k_stepNx<2>, then 3 + 1
    ref = oracle.c.run(*ref, 1.1, oracle.c.periodic(), 4)
    for a, b, nm in zip(lat.fields(), ref, 'f rho u'.split()):
        assert_parity(a, b, f'omega change after a two-step pass {nm}')
    lat.close()


def _karman_bundle(P, shape, second_plate=False):
    """milestone_6's rule set (inlet row, outlet rows, thin plate at nx//4) scaled to `shape`, as bench.py builds it."""
    B, BU = P.boundary_conditions, P.boundary_utils
    nx, ny = shape
    d = int(ny / 4.5) // 2 * 2
    plate = np.zeros(shape, dtype=bool)
    plate[nx // 4, ny // 2 - d // 2:ny // 2 + d // 2] = True
    bundle = BU.BoundaryBundle('von_karman_serial', shape)
    bundle.add(B.inlet(shape, 1.0, 0.1)).add(B.outlet()).add(B.rigid_object(plate))
    if second_plate:
        plate2 = np.zeros(shape, dtype=bool)
        plate2[5 * nx // 8, ny // 8:ny // 8 + 40] = True
        bundle.add(B.rigid_object(plate2))
    return bundle, d


@pytest.mark.parametrize('depth', [2, 3])
@pytest.mark.parametrize('shape,where,second', [((1024, 1024), 'clean', False), ((1024, 1024), 'plate', False),
                                                ((2050, 512), 'outlet', False), ((2050, 512), 'inlet', True),
                                                ((1024, 1024), 'next_to_strip', True)])
def test_two_steps_per_pass_with_boundary_cells(P, oracle, shape, where, second, depth):
    """Lattices WITH boundary cells take two steps per pass too: k_step2x on the rows whose two-step cone is all
    fluid, two one-step mask launches through a window on each strip of other rows (plan_strips). Must equal
    one-step launches and the C oracle bit for bit, probe ring included, wherever the probe sits."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    nx, ny = shape
    bundle, d = _karman_bundle(P, shape, second)
    f, rho, u = random_state(oracle, shape, 5)
    px, py = {'clean': (nx // 2, ny // 3), 'plate': (nx // 4 + 1, ny // 2), 'outlet': (nx - 2, 5), 'inlet': (1, ny - 1),
              'next_to_strip': (nx // 4 - 3, ny // 2 + 1)}[where]
    fused = Lattice(nx, ny, bundle.kind_map(shape))
    fused.set_option('fused_depth', depth)
    plain = Lattice(nx, ny, bundle.kind_map(shape))
    plain.set_option('fused', 0)
    for lat in (fused, plain):
        lat.probe(px, py, capacity=64)
        lat.load(f, rho, u, 1.41)
    l0, p0 = fused.launches, plain.launches
    for n in (7, 1, 8, 2):
        fused.run(n)
        plain.run(n)
    strips = 3 if second else 2          # {outlet rows, inlet row} wrap into one strip; one strip per plate
    clean = 2 if second else 1           # launches of the multi-step kernel over the clean ranges (two ranges per launch)
    assert plain.launches - p0 == 2 * 18
    # a pass of d steps = the multi-step kernel over the clean ranges + d mask launches per strip; a one-step launch = mask-free
    # kernel + edge list. Calls may END on a pass (results: strip windows + FINAL re-run of the clean rows).
    pp = lambda d: clean + d * strips
    if depth == 2:                        # 7 = 2+2+2+1, 1, 8 = 2+2+2+2, 2
        expect = (3 * pp(2) + 2) + 2 + 4 * pp(2) + pp(2)
    else:                                 # 7 = 3+3+1, 1, 8 = 3+3+2, 2
        expect = (2 * pp(3) + 2) + 2 + (2 * pp(3) + pp(2)) + pp(2)
    assert fused.launches - l0 == expect, 'the multi-step pass was not used'
    for a, b, nm in zip(fused.fields(), plain.fields(), 'f rho u'.split()):
        assert_parity(a, b, f'{shape} {where} {nm}')
    assert_parity(fused.probe_read(1, 18), plain.probe_read(1, 18), 'probe ring')
    if not second:
        scen = oracle.c.karman(nx, ny, 1.0, 0.1, d, ghost=0, probe=(px, py))
        ref = oracle.c.run(f, rho, u, 1.41, scen, 18, want_probe=True)
        for a, b, nm in zip(fused.fields(), ref, 'f rho u'.split()):
            assert_parity(a, b, f'vs oracle {nm}')
        assert_parity(fused.probe_read(1, 18), ref[3], 'probe vs oracle')
        sub = fused.fields(region=(nx // 4 - 3, nx // 4 + 4, ny // 2 - 5, ny // 2 + 6))     # a rectangle across strip and clean rows
        for a, b, nm in zip(sub, ref, 'f rho u'.split()):
            assert_parity(a, b[nx // 4 - 3:nx // 4 + 4, ny // 2 - 5:ny // 2 + 6], f'region {nm}')
        fused.run(4, 0.9)           # omega change right after a call that ended on a pass: its last collision is redone
        ref2 = oracle.c.run(ref[0], ref[1], ref[2], 0.9, scen, 4)
        for a, b, nm in zip(fused.fields(), ref2, 'f rho u'.split()):
            assert_parity(a, b, f'after omega change {nm}')
    fused.close()
    plain.close()


def test_boundary_rows_everywhere_stay_on_one_step_per_pass(P, oracle):
    """Walls along x make every row a boundary row: no clean rows, the two-step pass must not be chosen."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    shape = (1024, 1024)
    bc = P.boundary_utils.couette_flow_boundary_conditions(*shape, 0.03, 1.0)
    f, rho, u = random_state(oracle, shape, 6)
    lat = Lattice(*shape, bc.kind_map(shape))
    lat.load(f, rho, u, 0.8)
    l0 = lat.launches
    lat.run(6)
    assert lat.launches - l0 == 12       # mask-free kernel + edge list per step
    ref = oracle.c.run(f, rho, u, 0.8, oracle.c.couette(0.03, 1.0), 6)
    for a, b, nm in zip(lat.fields(), ref, 'f rho u'.split()):
        assert_parity(a, b, f'couette 1024 {nm}')
    lat.close()


# ---------------------------------------------------------------------------------------------------------
# cluster kernel: lattices that fit a thread-block cluster's distributed shared memory take all steps of a call in one launch
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case', ['periodic_100x50', 'couette_100x100', 'poiseuille_100x50', 'periodic_37x23', 'karman_64x48',
                                  'periodic_16x1000'])
def test_cluster_kernel_equals_single_launches_and_the_oracle(P, oracle, case):
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    BU = P.boundary_utils
    name, dims = case.split('_')
    shape = tuple(int(v) for v in dims.split('x'))
    omega, km, scen = 1.3, None, oracle.c.periodic()
    if name == 'couette':
        km, scen = BU.couette_flow_boundary_conditions(*shape, 0.05, 1.0).kind_map(shape), oracle.c.couette(0.05, 1.0)
    elif name == 'poiseuille':
        km, scen = BU.poiseuille_flow_boundary_conditions(*shape, 0.3338, 0.3328).kind_map(shape), oracle.c.poiseuille(0.3338, 0.3328)
    elif name == 'karman':
        bundle, d = _karman_bundle(P, shape)
        km, scen = bundle.kind_map(shape), oracle.c.karman(*shape, 1.0, 0.1, d, ghost=0, probe=(shape[0] // 2, 5))
    f, rho, u = random_state(oracle, shape, 13)
    px, py = shape[0] // 2, 5
    clu, one = Lattice(*shape, km), Lattice(*shape, km)
    clu.set_option('cluster', 2)            # always (the default times it against graph replay first and keeps the faster)
    one.set_option('cluster', 0)
    one.set_option('graphs', 0)
    for lat in (clu, one):
        lat.probe(px, py, capacity=256)
        lat.load(f, rho, u, omega)
    l0 = clu.launches
    for n in (7, 1, 40, 2, 33):            # odd and even launch lengths, a lone step in between
        clu.run(n)
        one.run(n)
    assert clu.launches - l0 == 5, 'one cluster launch per call (and a one-step launch) expected'
    for a, b, nm in zip(clu.fields(), one.fields(), 'f rho u'.split()):
        assert_parity(a, b, f'{case} {nm}')
    assert_parity(clu.probe_read(1, 83), one.probe_read(1, 83), 'probe ring')
    ref = oracle.c.run(f, rho, u, omega, scen, 83)
    for a, b, nm in zip(clu.fields(), ref, 'f rho u'.split()):
        assert_parity(a, b, f'{case} vs oracle {nm}')
    clu.run(10, 0.8)                        # omega change: the last step of the previous launch is redone
    ref = oracle.c.run(*ref, 0.8, scen, 10)
    for a, b, nm in zip(clu.fields(), ref, 'f rho u'.split()):
        assert_parity(a, b, f'{case} after omega change {nm}')
    # default policy: the first four long calls alternate between the cluster kernel and graph replay (timed), then one
    # of them is kept; whatever is chosen, the bits are the same
    auto = Lattice(*shape, km)
    auto.load(f, rho, u, omega)
    for _ in range(7):
        auto.run(64)
    ref = oracle.c.run(f, rho, u, omega, scen, 7 * 64)
    for a, b, nm in zip(auto.fields(), ref, 'f rho u'.split()):
        assert_parity(a, b, f'{case} auto-tuned path {nm}')
    auto.close()
    clu.close()
    one.close()


# ---------------------------------------------------------------------------------------------------------
# the whole job from and to host memory, pipelined over row chunks (lbm_run_host)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('steps,chunk,depth', [(20, 160, 3), (7, 96, 3), (6, 64, 3), (5, 200, 2), (13, 48, 4), (1, 64, 3)])
def test_streamed_run_from_and_to_host_memory(P, oracle, steps, chunk, depth):
    """Upload, time-skewed passes and download overlap chunk by chunk; results, probe samples and the state the context is
    left in must be those of load + run + fields — against the C oracle on a random field, for every remainder of the pass
    plan (ends on a three-step / two-step pass or a one-step launch) and with outputs aliasing the inputs."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    shape = (4096, 512)
    f, rho, u = random_state(oracle, shape, 100 + steps)
    lat = Lattice(*shape)
    lat.set_option('fused_depth', depth)
    lat.set_option('streamed_chunk_rows', chunk)
    px, py = 3, 77                       # a probe row inside the periodic seam
    lat.probe(px, py, capacity=64)
    l0 = lat.launches
    fo, ro, uo = f.copy(), rho.copy(), u.copy()
    out = lat.run_host(fo, ro, uo, 1.23, steps, out=(fo, ro, uo))        # in place
    assert lat.launches - l0 > 4096 // chunk, 'the streamed schedule was not taken'
    ref = oracle.c.run(f, rho, u, 1.23, oracle.c.periodic(), steps)
    for a, b, nm in zip(out, ref, 'f rho u'.split()):
        assert_parity(a, b, f'streamed {steps} steps {nm}')
    plain = Lattice(*shape)
    plain.set_option('fused', 0)
    plain.probe(px, py, capacity=64)
    plain.load(f, rho, u, 1.23)
    plain.run(steps)
    assert_parity(lat.probe_read(1, steps), plain.probe_read(1, steps), 'probe ring')
    for a, b, nm in zip(lat.fields(), ref, 'f rho u'.split()):          # the context holds the state of time `steps` ...
        assert_parity(a, b, f'state after the streamed run {nm}')
    lat.run(4, 0.9)                                                      # ... and goes on from it (with a new omega)
    plain.run(4, 0.9)
    for a, b, nm in zip(lat.fields(), plain.fields(), 'f rho u'.split()):
        assert_parity(a, b, f'continued {nm}')
    # not pipelined (option off): same answer through the three calls
    lat.set_option('streamed', 0)
    out2 = lat.run_host(f, rho, u, 1.23, steps)
    for a, b, nm in zip(out2, ref, 'f rho u'.split()):
        assert_parity(a, b, f'unstreamed {nm}')
    lat.close()
    plain.close()


@pytest.mark.parametrize('case', ['periodic_300x200', 'couette_100x100', 'karman_420x180', 'periodic_2048x1500'])
def test_kept_results_are_parked_in_the_device_history(P, oracle, case):
    """velocities.append(velocity) after every step (experiments.py:254, :542): kept density / velocity handles go to the
    device history (one asynchronous launch each, lbm_history_*), come back bit-exact when read — in any order, also after
    the lattice has been loaded again — and the last lattice is too large for slots (host materialisation instead)."""
    L, BU = P.lattice_boltzmann_method, P.boundary_utils
    from lattice_boltzmann_parallel_solver_b200 import engine
    name, dims = case.split('_')
    shape = tuple(int(v) for v in dims.split('x'))
    steps = 7 if shape[0] > 1000 else 40
    f, rho, u = random_state(oracle, shape, 11)
    if name == 'periodic':
        bundle, scen, omega = None, oracle.c.periodic(), 1.3
    elif name == 'couette':
        bundle, scen, omega = BU.couette_flow_boundary_conditions(*shape, 0.05, 1.0), oracle.c.couette(0.05, 1.0), 1.0
    else:
        import lattice_boltzmann_parallel_solver_b200 as pkg
        B = pkg.boundary_conditions
        plate = np.zeros(shape, dtype=bool); plate[shape[0] // 4, 70:110] = True
        bundle = BU.BoundaryBundle('von_karman_serial', shape)
        bundle.add(B.inlet(shape, 1.0, 0.1)).add(B.outlet()).add(B.rigid_object(plate))
        scen, omega = oracle.c.karman(shape[0], shape[1], 1.0, 0.1, 40, ghost=0), 1.6
    want, state = [], (f, rho, u)
    for _ in range(steps):
        state = oracle.c.run(*state, omega, scen, 1)
        want.append(state)
    kept_u, kept_rho = [], []
    for t in range(steps):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, omega, bundle)
        kept_u.append(u)
        if t % 3 == 0:
            kept_rho.append(rho)
    assert np.array_equal(np.asarray(f), want[-1][0])
    lat = kept_u[0]._lattice
    parked = [h for h in kept_u[:-1] if h._hist is not None]
    if shape[0] > 1000:
        assert not parked and lat._hist_free == []                   # 74 MB per slot: under four slots, history off
    else:
        assert len(parked) == steps - 1 and len(lat._hist_free) == min(engine.HISTORY_MAX_SLOTS, engine.HISTORY_BYTES // (shape[0] * shape[1] * 24)) - (steps - 1)
    order = np.random.default_rng(3).permutation(steps)
    for t in order[:steps // 2]:
        assert np.array_equal(np.asarray(kept_u[t]), want[t][2]), t
    # a fresh upload on the same lattice: parked results stay readable
    f2, rho2, u2 = L.lattice_boltzmann_step(*random_state(oracle, shape, 12), omega, bundle)
    np.asarray(rho2)
    assert kept_u[0]._lattice is rho2._lattice
    for t in order[steps // 2:]:
        assert np.array_equal(np.asarray(kept_u[t]), want[t][2]), t
    for i, h in enumerate(kept_rho):
        assert np.array_equal(np.asarray(h), want[3 * i][1]), i
    L.release_lattices()                                             # parked-but-unread handles survive the lattice
    assert all(h._value is not None for h in kept_u + kept_rho)


def test_options_and_state_errors(P, oracle):
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    lat = Lattice(32, 32)
    with pytest.raises(AssertionError, match='unknown option'):
        lat.set_option('warp_drive', 1)
    with pytest.raises(N.LbmStateError):
        lat.run(1, 1.0)                                  # nothing loaded
    f, rho, u = random_state(oracle, (32, 32), 1)
    lat.load(f, rho, u, 1.0)
    with pytest.raises(N.LbmStateError):
        lat.fields()                                     # no step since the load: the caller still holds the state
    with pytest.raises(N.LbmStateError):
        lat.run(1, 1.5)                                  # omega differs from the load's and nothing to redo from
    lat.run(2, 1.0)
    with pytest.raises(AssertionError):
        lat.fields(region=(0, 33, 0, 32))
    with pytest.raises(AssertionError):
        lat.probe(40, 0)
    lat.close()
    with pytest.raises(AssertionError):
        Lattice(0, 5)
    with pytest.raises(AssertionError):
        Lattice(64, 64, ghost=(2, 1))
    with pytest.raises(AssertionError):
        Lattice(64, 64, ghost=(5, 0))
