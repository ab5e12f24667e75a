"""CPU-only tests of the host side: C-ABI surface, topology math, boundary compilation, the process-group
communicator over gloo (world_size 2 and 4), failure modes without a GPU."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def test_library_exports_every_declared_symbol():
    """include/lbm_b200.h <-> liblbm_b200.so <-> the ctypes table, no compute call."""
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    header = open(os.path.join(ROOT, 'include', 'lbm_b200.h')).read()
    declared = set(re.findall(r'\b(lbm_[a-z_0-9]+)\s*\(', header))
    declared -= {'lbm_ctx', 'lbm_kind', 'lbm_bc_desc', 'lbm_halo_export'}
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    lib = N.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b'sm_100a' in lib.lbm_version()
    assert ctypes.sizeof(N.Kind) == 12 and ctypes.sizeof(N.HaloExport) == 104


def test_no_gpu_means_loud_failure():
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    if N.load().lbm_device_count() > 0:
        pytest.skip('a GPU is visible')
    import lattice_boltzmann_parallel_solver_b200.lattice_boltzmann_method as L
    with pytest.raises(N.LbmNativeError):
        L.compute_density(np.ones((4, 4, 9)))
    with pytest.raises(N.LbmNativeError):
        L.lattice_boltzmann_step(np.ones((4, 4, 9)), np.ones((4, 4)), np.zeros((4, 4, 2)), 1.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'lattice_boltzmann_parallel_solver_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), fn
                assert 'liblbm_oracle' not in src, fn


def test_topology_vs_reference_tables():
    from lattice_boltzmann_parallel_solver_b200 import parallelization_utils as PU
    with open(os.path.join(GOLDEN, 'topology.json')) as fh:
        g = json.load(fh)
    for n, v in g['get_xy_size'].items():
        if v is None:
            with pytest.raises(Exception, match='prime'):
                PU.get_xy_size(int(n))
        else:
            xs, ys = PU.get_xy_size(int(n))
            assert (float(xs), float(ys)) == (v[0], v[1]) and type(xs).__name__ == v[2]
    for n, lx, ly, cx, cy, nlx, nly in g['local_coords']:
        xs, ys = PU.get_xy_size(n)
        assert PU.get_local_coords([cx, cy], lx, ly, xs, ys) == (nlx, nly)
    for n, lx, ly, cx, cy, gx, gy, xin, yin, lxx, lyy, loc in g['in_process']:
        xs, ys = PU.get_xy_size(n)
        c = [cx, cy]
        assert bool(PU.x_in_process(c, gx, lx, xs)) == xin and bool(PU.y_in_process(c, gy, ly, ys)) == yin
        assert PU.global_to_local_direction(cx, gx, lx, xs) == lxx
        assert PU.global_to_local_direction(cy, gy, ly, ys) == lyy
        r = PU.global_coord_to_local_coord(c, gx, gy, lx, ly, xs, ys)
        assert (None if r[0] is None else [r[1], r[2]]) == loc
    from lattice_boltzmann_parallel_solver_b200 import lattice_boltzmann_method as L
    assert float(L.reynolds_number(40, 0.1, 0.04)) == g['reynolds']
    assert float(L.strouhal_number(1.1308e-3, 40, 0.1)) == g['strouhal']
    assert np.array_equal(L.get_velocity_sets(), np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1]]))
    assert np.array_equal(L.vel_to_opp_vel_mapping(), [0, 3, 4, 1, 2, 7, 8, 5, 6])
    assert np.array_equal(L.get_w_i(), [4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)


def _rules_from_oracle(bc_factory, shape, n_args):
    """Derive the per-slot overwrite table of an ORACLE boundary closure by probing it with marker arrays, to
    compare with the kind map this package compiles for the same scenario."""
    nx, ny = shape
    idx = np.arange(nx * ny * 9, dtype=np.float64).reshape(nx, ny, 9)
    pre, post, prev = idx + 1e7, idx + 2e7, idx + 3e7
    out = bc_factory(pre.copy(), post.copy(), np.ones(shape), np.zeros(shape + (2,)), prev.copy())
    return out, pre, post, prev


def test_boundary_bundles_compile_to_the_oracles_overwrites():
    """Without a GPU: the compiled kind map must describe exactly the overwrites the oracle's closures perform
    (which slots, from where), for Couette, the serial plate and every rank of the 2x3 and 2x7 parallel layouts."""
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200 import boundary_spec as S
    from lattice_boltzmann_parallel_solver_b200 import boundary_utils as BU
    from lattice_boltzmann_parallel_solver_b200 import parallelization_utils as PU
    from lattice_boltzmann_parallel_solver_b200.boundary_conditions import BoundaryOp
    from oracle import lbm_numpy as onp
    # inlet constants come from the GPU in the product; patch them with the oracle's for this CPU-only test
    import lattice_boltzmann_parallel_solver_b200.boundary_conditions as B

    def check(km, oracle_bc, shape):
        out, pre, post, prev = _rules_from_oracle(oracle_bc, shape, 5)
        t = km.table
        for x in range(shape[0]):
            for y in range(shape[1]):
                rules, flags, skip = t.kinds[int(km.map[x, y])]
                for i in range(9):
                    typ, row = rules[i] & 7, rules[i] >> 3
                    got = out[x, y, i]
                    if typ == N.RULE_PULL:
                        assert got == post[x, y, i], (x, y, i)
                    elif typ == N.RULE_BOUNCE:
                        assert got == pre[x, y, S.OPP[i]] - t.k_rows[row][S.OPP[i]], (x, y, i)
                    elif typ == N.RULE_OUTLET:
                        assert got == prev[x - 1, y, i], (x, y, i)
                    else:
                        assert got == t.c_rows[row][i], (x, y, i)

    real_eq = B.equilibrium_distr_func
    B.equilibrium_distr_func = lambda rho, u: onp.equilibrium(rho, u)
    try:
        shape = (12, 9)
        check(BU.couette_flow_boundary_conditions(*shape, 0.05, 1.0).kind_map(shape), onp.couette_bc(*shape, 0.05, 1.0), shape)
        lx, ly, d = 40, 30, 8
        for size in (1, 6, 14):
            xs, ys = PU.get_xy_size(size)
            for cx in range(int(xs)):
                for cy in range(int(ys)):
                    c = [cx, cy]
                    nlx, nly = PU.get_local_coords(c, lx, ly, xs, ys)
                    shp = (nlx + 2, nly + 2)
                    km = BU.parallel_von_karman_boundary_conditions(c, nlx, nly, lx, ly, xs, ys, 1.0, 0.1, d).kind_map(shp)
                    check(km, onp.karman_parallel_bc(c, nlx, nly, lx, ly, int(xs), int(ys), 1.0, 0.1, d), shp)
                    assert not km.map[0].any() and not km.map[-1].any() and not km.map[:, 0].any() and not km.map[:, -1].any()
        # serial plate through rigid_object
        shape = (24, 20)
        plate = np.zeros(shape, dtype=bool)
        plate[6, 6:14] = True
        b = BU.BoundaryBundle('m6', shape).add(B.inlet(shape, 1.0, 0.1)).add(B.outlet()).add(B.rigid_object(plate))
        check(b.kind_map(shape), onp.karman_serial_bc(24, 20, 1.0, 0.1, 8), shape)
        # the split outlet is refused exactly like the reference (boundary_utils.py:174-176)
        with pytest.raises(NotImplementedError):
            BU.parallel_von_karman_boundary_conditions([20, 0], 1, 30, 21, 30, 21, 1, 1.0, 0.1, 8)
        # Poiseuille: flags on rows 1 / -2, skipped stores on the virtual rows, bounce on both walls
        km = BU.poiseuille_flow_boundary_conditions(10, 6, 0.3345, 0.3321).kind_map((10, 6))
        t = km.table
        assert t.rho_in == float(np.divide(0.3345, onp.CS2)) and t.rho_out == float(np.divide(0.3321, onp.CS2))
        for y in range(6):
            assert t.kinds[km.map[8, y]][1] & N.CELL_PBC_IN_SRC and t.kinds[km.map[1, y]][1] & N.CELL_PBC_OUT_SRC
            assert t.kinds[km.map[0, y]][2] == (1 << 1) | (1 << 5) | (1 << 8)
            assert t.kinds[km.map[9, y]][2] == (1 << 3) | (1 << 6) | (1 << 7)
        for x in range(10):
            assert [r & 7 for r in t.kinds[km.map[x, 0]][0]][2] == N.RULE_BOUNCE      # f_post[.,2] <- f_pre[.,4] at y=0
            assert [r & 7 for r in t.kinds[km.map[x, 5]][0]][4] == N.RULE_BOUNCE
        assert isinstance(b.ops[0][0], BoundaryOp)
    finally:
        B.equilibrium_distr_func = real_eq


def test_unknown_boundary_callable_is_rejected():
    from lattice_boltzmann_parallel_solver_b200 import lattice_boltzmann_method as L
    with pytest.raises(TypeError, match='no CPU fallback'):
        L._resolve_boundary(lambda *a: a[1], (4, 4))


def test_pressure_periodic_y_variant_is_refused():
    from lattice_boltzmann_parallel_solver_b200 import boundary_conditions as B
    m = np.zeros((6, 6), dtype=bool)
    m[:, 0] = True
    m[:, -1] = True
    m[0, 1] = True
    op = B.periodic_with_pressure_variations(m, 0.3345, 0.3321)
    with pytest.raises(NotImplementedError):
        op(np.ones((6, 6, 9)), np.ones((6, 6)), np.zeros((6, 6, 2)))
    with pytest.raises(AssertionError):
        B.rigid_wall(np.zeros((4, 4)))           # non-bool mask, boundary_conditions.py:89


@pytest.mark.parametrize('size', [2, 4])
def test_gloo_ranks_host_path(size):
    """world_size > 1 over gloo on the CPU: CartComm (coords/Shift/Sendrecv), communication(f) on host arrays,
    the per-rank bundles' geometry and the save_mpiio gather, against the reference's own k-rank results."""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={size}',
           '--master-addr', '127.0.0.1', '--master-port', str(29620 + size), os.path.join(ROOT, 'tests', 'mp_karman.py'),
           '--host']
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', OMP_NUM_THREADS='1')
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert f'OK {size} ranks (host/gloo)' in res.stdout


def test_cart_comm_matches_mpi_cart_semantics():
    from lattice_boltzmann_parallel_solver_b200 import dist
    w = dist.WorldComm()
    assert w.Get_size() == 1 and w.Get_rank() == 0
    cart = w.Create_cart(dims=[1, 1], periods=[True, True], reorder=False)
    assert cart.Get_coords(0) == [0, 0] and cart.Shift(0, 1) == (0, 0) and cart.Shift(1, -1) == (0, 0)
    f = np.arange(6 * 5 * 9, dtype=np.float64).reshape(6, 5, 9)
    from lattice_boltzmann_parallel_solver_b200 import parallelization_utils as PU
    from oracle import lbm_numpy as onp
    assert np.array_equal(PU.communication(cart)(f.copy()), onp.self_exchange(f.copy()))
    with pytest.raises(AssertionError):
        w.Create_cart(dims=[2, 1])


def test_dropin_modules_expose_the_reference_api():
    """Flat imports from the dropin directory (how the reference's drivers import, Makefile:3) expose every public
    function of the reference modules with the same parameter names. Names are pinned here; when the reference
    tree is present (build container) the list is re-derived from it."""
    import importlib
    import inspect
    dropin = os.path.join(ROOT, 'lattice_boltzmann_parallel_solver_b200', 'dropin')
    sys.path.insert(0, dropin)
    try:
        expected = {
            'lattice_boltzmann_method': {
                'get_velocity_sets': [], 'vel_to_opp_vel_mapping': [], 'get_w_i': [], 'reynolds_number': ['L', 'u', 'v'],
                'strouhal_number': ['f', 'L', 'u'], 'compute_density': ['prob_densitiy_func'],
                'compute_velocity_field': ['density_func', 'prob_density_func'], 'streaming': ['prob_density_func'],
                'equilibrium_distr_func': ['density_func', 'velocity_field'],
                'lattice_boltzmann_step': ['f', 'density', 'velocity', 'omega', 'boundary', 'parallel_communication']},
            'boundary_conditions': {
                'get_wall_indices': ['boundary'], 'get_corner_indices': ['boundary'],
                'remove_corner_indices_from_boundary': ['boundary', 'corner_indices'], 'rigid_wall': ['boundary'],
                'rigid_object': ['boundary'], 'moving_wall': ['boundary', 'u_w', 'avg_density'],
                'inlet': ['lattice_grid_shape', 'density_in', 'velocity_in'], 'outlet': [],
                'periodic_with_pressure_variations': ['boundary', 'p_in', 'p_out']},
            'boundary_utils': {
                'couette_flow_boundary_conditions': ['lx', 'ly', 'U', 'avg_density'],
                'poiseuille_flow_boundary_conditions': ['lx', 'ly', 'p_in', 'p_out'],
                'parallel_von_karman_boundary_conditions': ['coord2d', 'n_local_x', 'n_local_y', 'lx', 'ly', 'x_size',
                                                            'y_size', 'density_in', 'velocity_in', 'plate_size']},
            'parallelization_utils': {
                'communication': ['comm'], 'get_xy_size': ['total_number_of_procces'],
                'get_local_coords': ['coords2d', 'lx', 'ly', 'x_size', 'y_size'],
                'global_to_local_direction': ['coord1d', 'global_dir', 'lattice_dir', 'dir_size'],
                'global_coord_to_local_coord': ['coord2d', 'global_x', 'global_y', 'lx', 'ly', 'x_size', 'y_size'],
                'x_in_process': ['coord2d', 'x_coord', 'lx', 'processes_in_x'],
                'y_in_process': ['coord2d', 'y_coord', 'ly', 'processes_in_y'], 'save_mpiio': ['comm', 'fn', 'g_kl']},
        }
        ref_src = '/root/reference/src'
        if os.path.isdir(ref_src):
            import ast
            for mod, table in expected.items():
                tree = ast.parse(open(os.path.join(ref_src, mod + '.py')).read())
                found = {n.name: [a.arg for a in n.args.args] for n in tree.body if isinstance(n, ast.FunctionDef)}
                assert found == table, mod
        for mod, table in expected.items():
            for cached in [m for m in sys.modules if m == mod]:
                del sys.modules[cached]
            m = importlib.import_module(mod)
            assert m.__file__.startswith(dropin)
            for name, params in table.items():
                fn = getattr(m, name)
                assert list(inspect.signature(fn).parameters) == params, (mod, name)
        from mpi4py import MPI
        assert MPI.COMM_WORLD.Get_size() == 1 and hasattr(MPI, 'Intracomm')
    finally:
        sys.path.remove(dropin)
        for mod in list(expected) + ['mpi4py', 'mpi4py.MPI']:
            sys.modules.pop(mod, None)


def test_cpu_baseline_processes_equal_one_process():
    """bench.py's all-cores CPU arm (k processes, slabs, shared-memory ghost exchange; the reference's own time step
    where baseline/_ref/src exists, else the oracle port) computes exactly what the single-process oracle computes."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import lbm_numpy as onp
    n, steps = 64, 5
    _, _, got = bench.cpu_reference_mlups_parallel(n, steps, 0, 4, return_fields=True)
    rho, u = onp.sinusoidal_velocity_x((n, n), bench.EPS)
    f = onp.equilibrium(rho, u)
    for _ in range(steps):
        f, rho, u = onp.step(f, rho, u, bench.OMEGA)
    assert np.array_equal(got, f)
    line = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup',
                           '1', '--cpu-size', '256'], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert line.returncode == 0, line.stderr[-2000:]
    assert len(line.stdout.strip().splitlines()) == 1, 'bench.py must print exactly one line on stdout'
    d = json.loads(line.stdout.strip().splitlines()[-1])
    have_ref = os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'src', 'lattice_boltzmann_method.py'))
    assert d['impl'] == 'reference' and d['value'] > 0 and d['e2e']['value'] == d['value']
    assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port')      # the unmodified reference where it travelled
    assert d['config']['lattice'] == [256, 256] and '256x256' in d['config']['workload']   # states its OWN workload


# ---------------------------------------------------------------------------------------------------------
# row plan of the two-steps-per-pass schedule on lattices with boundary cells (host logic of the C library)
# ---------------------------------------------------------------------------------------------------------
def _plan(nx, dirty_rows):
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    lib = N.load()
    flags = (ctypes.c_uint8 * nx)()
    for r in dirty_rows:
        flags[r] = 1
    ns, nc = ctypes.c_int(), ctypes.c_int()
    strips, clean = (ctypes.c_int * 32)(), (ctypes.c_int * 32)()
    N.check(lib.lbm_plan_two_step(nx, flags, ctypes.byref(ns), strips, ctypes.byref(nc), clean))
    if ns.value < 0:
        return None
    return ([(strips[2 * i], strips[2 * i + 1]) for i in range(ns.value)],
            [(clean[2 * i], clean[2 * i + 1]) for i in range(nc.value)])


def test_two_step_row_plan_known_cases():
    nx = 16384
    # von Karman rule set: inlet row 0, plate rows nx/4 and nx/4+1, outlet rows nx-2, nx-1 (bench.py karman_lattice)
    strips, clean = _plan(nx, [0, 4096, 4097, nx - 2, nx - 1])
    assert strips == [(4094, 4100), (nx - 4, nx + 3)]          # the outlet / inlet rows wrap into ONE strip
    assert clean == [(3, 4094), (4100, nx - 4)]
    assert _plan(nx, []) == ([], [(0, nx)])
    assert _plan(nx, [8000]) == ([(7998, 8003)], [(0, 7998), (8003, nx)])
    assert _plan(nx, [1]) == ([(nx - 1, nx + 4)], [(4, nx - 1)])
    assert _plan(64, range(64)) is None                          # walls along x: every row is a boundary row
    assert _plan(64, range(0, 64, 5)) is None                    # a boundary row within two rows of every row
    assert _plan(8, [3]) is None                                 # too few rows
    assert _plan(4096, range(0, 4096, 128)) is None              # 32 strips: more ranges than the schedule keeps


def test_two_step_row_plan_properties():
    """Every row is in exactly one strip or clean range; a clean row has no boundary row within two rows (periodic);
    strips are maximal runs; ranges are sorted."""
    rng = np.random.default_rng(4)
    for trial in range(200):
        nx = int(rng.integers(16, 400))
        dirty = sorted(set(int(v) for v in rng.integers(0, nx, int(rng.integers(0, 6)))))
        plan = _plan(nx, dirty)
        near = np.zeros(nx, dtype=bool)
        for r in dirty:
            for d in range(-2, 3):
                near[(r + d) % nx] = True
        if plan is None:
            assert 2 * near.sum() > nx, (nx, dirty)
            continue
        strips, clean = plan
        owner = np.zeros(nx, dtype=int)
        for a, b in strips:
            assert 0 <= a < nx and a < b
            for r in range(a, b):
                owner[r % nx] += 1
                assert near[r % nx]
            assert not near[(a - 1) % nx] and not near[b % nx]   # maximal
        for a, b in clean:
            assert 0 <= a < b <= nx
            owner[a:b] += 1
            assert not near[a:b].any()
        assert (owner == 1).all(), (nx, dirty, plan)
        assert clean == sorted(clean) and strips == sorted(strips)


def test_karman_slab_kind_maps_equal_the_global_construction(monkeypatch):
    """bench.py builds the von Karman rule set of an N-GPU lattice slab by slab (ghost rows included, which carry the
    neighbour's kinds) without the global arrays; cell by cell it must be the rule set the package's own operators
    (inlet, outlet, rigid_object — bench.karman_lattice) compile on the global lattice."""
    sys.path.insert(0, ROOT)
    import bench
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200 import boundary_conditions as B
    from lattice_boltzmann_parallel_solver_b200 import boundary_utils as BU
    from tests.fake_native import FakeLib
    fake = FakeLib()
    monkeypatch.setattr(N, 'load', lambda: fake)
    monkeypatch.setattr(N, 'device', lambda: 0)

    def effective(km):
        """per cell: (rule of each population with its K / constant row RESOLVED to values, flags, skip)"""
        out = {}
        t = km.table
        for k, (rules, flags, skip) in enumerate(t.kinds):
            res = []
            for i, r in enumerate(rules):
                typ, row = r & 7, r >> 3
                val = tuple(t.k_rows[row]) if typ == N.RULE_BOUNCE else (tuple(t.c_rows[row]) if typ == N.RULE_CONST else ())
                res.append((typ, val))
            out[k] = (tuple(res), flags, skip)
        return [[out[int(k)] for k in row] for row in km.map]

    for nxg, ny, world, g in ((64, 36, 4, 2), (96, 45, 3, 2), (40, 20, 2, 3)):
        d = int(ny / 4.5) // 2 * 2
        plate = np.zeros((nxg, ny), dtype=bool)
        plate[nxg // 4, ny // 2 - d // 2:ny // 2 + d // 2] = True
        bundle = BU.BoundaryBundle('von_karman_serial', (nxg, ny))
        bundle.add(B.inlet((nxg, ny), 1.0, 0.1)).add(B.outlet()).add(B.rigid_object(plate))
        want = effective(bundle.kind_map((nxg, ny)))
        nxl = nxg // world
        for rank in range(world):
            got = effective(bench.karman_slab_kind_map(nxg, ny, rank * nxl - g, nxl + 2 * g))
            for i in range(nxl + 2 * g):
                assert got[i] == want[(rank * nxl - g + i) % nxg], (nxg, ny, world, rank, i)


def test_clock_sampler_reports_the_rows_of_the_timed_region(tmp_path, monkeypatch):
    """bench.ClockSampler against a stand-in nvidia-smi: rows that arrive between begin() and end() are the ones reported
    (throttle reasons parsed); a region shorter than a poll falls back to the rows since the warm-up began."""
    import stat
    import time
    import bench
    fake = tmp_path / 'nvidia-smi'
    fake.write_text('#!/bin/bash\n'
                    'i=0\n'
                    'while true; do\n'
                    '  if [ $i -lt 5 ]; then echo "0, 345, 1965, 200.0, 0x0, Not Active, Not Active, Not Active, Not Active";\n'
                    '  else echo "0, 1905, 1965, 990.0, 0x4, Not Active, Not Active, Not Active, Active"; fi\n'
                    '  i=$((i+1)); sleep 0.02\n'
                    'done\n')
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv('PATH', str(tmp_path) + os.pathsep + os.environ['PATH'])
    s = bench.ClockSampler(0)
    s.start()
    assert s.rows, 'the first row is waited for'
    time.sleep(0.15)          # "warm-up": the idle rows (345 MHz) pass
    s.begin()
    time.sleep(0.12)          # "timed region"
    s.end()
    got = s.stop()
    assert got['window'] == 'timed region' and got['samples'] >= 3, got
    assert got['sm_mhz'] == 1905.0 and got['sm_max_mhz'] == 1965.0 and got['reasons'] == ['sw_power_cap'], got
    class Done:                # a region no poll fell into (nvidia-smi slower than the region): rows since the warm-up began
        def terminate(self):
            pass
    s = bench.ClockSampler(0)
    s.proc, s.t_start, s.t_begin, s.t_end = Done(), 10.0, 12.0, 12.05
    s.rows = [(9.0, '0, 345, 1965, 200.0, 0x0, Not Active, Not Active, Not Active, Not Active'),
              (11.0, '0, 1935, 1965, 990.0, 0x0, Not Active, Not Active, Not Active, Not Active'),
              (13.0, '0, 345, 1965, 200.0, 0x0, Not Active, Not Active, Not Active, Not Active')]
    got = s.stop()
    assert got['samples'] == 1 and got['sm_mhz'] == 1935.0 and got['window'].startswith('warm-up') and got['reasons'] == [], got


def test_header_documents_every_option():
    """include/lbm_b200.h is the contract: every option name lbm_set_option accepts is documented there and listed in the
    error message for an unknown name."""
    import re
    src = open(os.path.join(ROOT, 'lattice_boltzmann_parallel_solver_b200', 'csrc', 'lbm_b200.cu')).read()
    body = src[src.index('extern "C" int lbm_set_option'):src.index('extern "C" int64_t lbm_device_bytes')]
    names = set(re.findall(r'n == "([a-z0-9_]+)"', body))
    assert len(names) >= 15, names
    header = open(os.path.join(ROOT, 'include', 'lbm_b200.h')).read()
    listed = body[body.index('unknown option'):]
    for name in sorted(names):
        assert f'"{name}"' in header, f'option {name} is not documented in include/lbm_b200.h'
        assert re.search(r'\b' + name + r'\b', listed), f'option {name} is missing from the unknown-option message'
