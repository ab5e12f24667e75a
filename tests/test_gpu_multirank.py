"""k-rank decompositions on the GPU: (a) all k blocks as k device lattices of ONE process (runs on a 1-GPU box;
exercises the ghost stores, the 8-neighbour wiring and the step-flag protocol between distinct contexts), and
(b) real one-process-per-GPU runs under torchrun with CUDA-IPC peer stores (needs >= 2 GPUs).
Reference design: tests/test_parallelization_von_karman.py under `mpirun -N {1,2,4,6,8,9,14}` (.travis.yml:20-26):
the parallel run must equal the serial one; here additionally every rank's whole local array, ghost ring
included, must equal what the reference itself produced on k ranks (tests/golden/karman_ranks.npz)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import sha

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
KARMAN = dict(lx=420, ly=180, d=40, u0=0.1, rho_in=1.0, nu=0.04)


def _gpu_count():
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    return N.load().lbm_device_count()


@pytest.mark.parametrize('size', [2, 4, 6, 8, 9, 14])
def test_k_blocks_in_one_process(size):
    import lattice_boltzmann_parallel_solver_b200 as P
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice, connect_blocks
    from oracle import lbm_numpy as onp
    PU, BU, L = P.parallelization_utils, P.boundary_utils, P.lattice_boltzmann_method
    g = np.load(os.path.join(GOLDEN, 'karman_ranks.npz'))
    k = KARMAN
    lx, ly = k['lx'], k['ly']
    omega = float(np.reciprocal(3 * k['nu'] + 0.5))
    x_size, y_size = PU.get_xy_size(size)
    xs, ys = int(x_size), int(y_size)
    ndev = _gpu_count()
    blocks, init = {}, {}
    for cx in range(xs):
        for cy in range(ys):
            c = [cx, cy]
            nlx, nly = PU.get_local_coords(c, lx, ly, x_size, y_size)
            bc = BU.parallel_von_karman_boundary_conditions(c, nlx, nly, lx, ly, x_size, y_size, k['rho_in'], k['u0'], k['d'])
            blocks[(cx, cy)] = Lattice(nlx + 2, nly + 2, bc.kind_map((nlx + 2, nly + 2)), ghost=(1, 1),
                                       device=(cx * ys + cy) % ndev)
            rho, u = onp.uniform((nlx + 2, nly + 2), 1.0, k['u0'], 0.0)
            init[(cx, cy)] = (L.equilibrium_distr_func(rho, u), rho, u)
    connect_blocks(blocks, (xs, ys))
    for c, lat in blocks.items():
        lat.load(*init[c], omega)
    for lat in blocks.values():
        lat.sync()
    from lattice_boltzmann_parallel_solver_b200.engine import run_blocks
    run_blocks(blocks, 11)             # lockstep launch groups, drained between groups when blocks share a device
    G = np.zeros((lx, ly, 9))
    for (cx, cy), lat in blocks.items():
        lat.sync()
        f, rho, u = lat.fields()
        r = cx * ys + cy
        assert sha(f) == str(g[f'n{size}_r{r}_f_full']), f'rank {r} f (ghost ring included)'
        assert sha(rho) == str(g[f'n{size}_r{r}_rho_full']), f'rank {r} rho'
        assert sha(u) == str(g[f'n{size}_r{r}_u_full']), f'rank {r} u'
        x0, y0 = cx * (lx // xs), cy * (ly // ys)
        G[x0:x0 + f.shape[0] - 2, y0:y0 + f.shape[1] - 2] = f[1:-1, 1:-1]
    assert sha(G) == str(g['serial_f11'])
    # A neighbour that runs one step ahead overwrites this block's ghost cells of the buffer time 11 is rebuilt
    # from; the ghost snapshot must keep the materialised values (ghost ring and first interior ring) intact.
    first = blocks[(0, 0)]
    before = first.fields()
    for c, lat in blocks.items():
        if lat is not first:
            lat.run(1)
    for c, lat in blocks.items():
        if lat is not first:
            lat.sync()
    after = first.fields()
    for a, b, nm in zip(before, after, 'f rho u'.split()):
        assert np.array_equal(a, b), f'time-11 {nm} of block (0,0) changed after its neighbours took step 12'
    for lat in blocks.values():
        lat.close()


def test_slabs_equal_single_block():
    """The bench decomposition (1-D slabs along x, ghost rows only, in-kernel periodic wrap along y, split edge /
    interior launches on two streams) must give exactly the single-block result."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice, connect_blocks
    from oracle import lbm_c, lbm_numpy as onp
    nx, ny, k, steps = 4096, 512, 4, 12          # 2M cells: above the threshold that enables the split launch
    rho, u = onp.sinusoidal_velocity_x((nx, ny), 0.01)
    rng = np.random.default_rng(2)
    rho = rho * rng.uniform(0.99, 1.01, rho.shape)    # break the x-invariance so that halo errors are visible
    f = onp.equilibrium(rho, u)
    ref = lbm_c.run(f, rho, u, 1.1, lbm_c.periodic(), steps)
    ndev = _gpu_count()
    n = nx // k
    blocks = {}
    for c in range(k):
        lat = Lattice(n + 2, ny, ghost=(1, 0), device=c % ndev)
        blocks[(c, 0)] = lat
    connect_blocks(blocks, (k, 1))

    def padded(a, c):
        idx = np.arange(c * n - 1, (c + 1) * n + 1) % nx
        return np.ascontiguousarray(a[idx])
    for c in range(k):
        blocks[(c, 0)].load(padded(f, c), padded(rho, c), padded(u, c), 1.1)
    for lat in blocks.values():
        lat.sync()
    for _ in range(steps):
        for lat in blocks.values():
            lat.run(1)
    for c in range(k):
        lat = blocks[(c, 0)]
        lat.sync()
        got = lat.fields(region=(1, n + 1, 0, ny))
        for a, b, nm in zip(got, ref, 'f rho u'.split()):
            assert np.array_equal(a, b[c * n:(c + 1) * n]), f'slab {c} {nm}'
        lat.close()


def test_two_row_slabs_two_steps_per_pass():
    """The bench decomposition since the two-steps-per-pass kernel: slabs with TWO ghost rows per side; edge rows
    (2 + 2) on the edge stream with ghost stores, interior on the main stream. Must equal the single block.
    Own process with CUDA_DEVICE_MAX_CONNECTIONS=32 (see tests/mp_slabs.py)."""
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS='32', LBM_HALO_TIMEOUT_S='10')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'mp_slabs.py'), '--inproc', '4'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'OK 4 slabs in one process' in res.stdout


def _torchrun(script, size, port, *args, timeout=900, halo_timeout='120'):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={size}',
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', script)] + list(args)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                         env=dict(os.environ, LBM_HALO_TIMEOUT_S=halo_timeout))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


def _placement(size):
    """One process per GPU where the box has `size` GPUs (NCCL plumbing); otherwise ALL ranks on GPU 0 (gloo plumbing):
    the same cudaIpcOpenMemHandle mappings, peer ghost stores from inside the step kernels and flag handshake between
    PROCESSES, with the ranks' kernels time-sliced on one device (a kernel that waits for another rank's flag is
    preempted; generous timeout). So the cross-process halo path is exercised on a single-GPU box too."""
    return [] if _gpu_count() >= size else ['--shared-gpu']


@pytest.mark.parametrize('size,depth', [(2, 2), (2, 3), (4, 3), (8, 3)])
def test_slabs_one_process_per_rank(size, depth):
    """The bench decomposition across processes: slabs with `depth` ghost rows, `depth` steps per pass, edge rows with
    peer ghost stores + flag handshake on the edge stream, interior on the main stream; calls that end on a pass
    (FINAL-mode materialisation) and on a one-step launch. Values must equal the single-block C oracle on a field
    that varies along the slab axis (tests/mp_slabs.py)."""
    place = _placement(size)
    out = _torchrun('mp_slabs.py', size, 29540 + size + depth, '--depth', str(depth), *place)
    assert f'OK {size} slabs, one process per GPU' + (' (shared)' if place else '') + f', depth {depth}' in out


@pytest.mark.parametrize('size', [2, 4, 8])
def test_karman_one_process_per_rank(size):
    """tests/test_parallelization_von_karman.py:18-66 of the reference across `size` processes on its own get_xy_size
    grid: every rank's whole local arrays (ghost ring included) must equal what the reference produced on as many
    ranks, the probe trace and the save_mpiio gather too (tests/mp_karman.py)."""
    place = _placement(size)
    out = _torchrun('mp_karman.py', size, 29500 + size, *place)
    assert f'OK {size} ranks ({"one gpu/gloo+ipc" if place else "gpu/nccl+ipc"})' in out


@pytest.mark.parametrize('size,depth', [(2, 3), (4, 3), (8, 3), (4, 2)])
def test_karman_slabs_multi_step_passes_across_ranks(size, depth):
    """Boundary-bearing lattices on N ranks take multi-step passes too (slabs with as many ghost rows that carry the
    neighbour's kinds; strip windows next to boundary rows and slab edges with ghost stores + flag handshake). With 4
    ranks the plate sits on the first row of rank 1, i.e. on rank 0's ghost rows; with 8 ranks rank 1 has boundary cells
    on its ghost rows ONLY. Must equal the single-block oracle."""
    place = _placement(size)
    out = _torchrun('mp_karman_slabs.py', size, 29580 + size + depth, '--depth', str(depth), *place)
    assert f'OK {size} karman slabs, {depth} steps per pass' in out


@pytest.mark.parametrize('size,depth', [(2, 3), (4, 3), (2, 2)])
def test_streamed_run_on_slabs_across_ranks(size, depth):
    """lbm_run_host on slabs: upload, time-skewed passes and download pipelined over row chunks on every rank; the rows
    near the slab edges are finished last, pass by pass with ghost stores and flag handshake (the upload phase publishes
    an epoch of its own, no host barrier inside the call). Interior rows must equal the single-block oracle, and the
    context must hold the state of the last step (check() materialises it again)."""
    place = _placement(size)
    out = _torchrun('mp_slabs.py', size, 29640 + size + depth, '--depth', str(depth), '--run-host', *place)
    assert f'OK {size} slabs, one process per GPU' + (' (shared)' if place else '') + f', depth {depth}, run_host' in out
