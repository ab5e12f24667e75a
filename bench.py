#!/usr/bin/env python
"""bench.py — D2Q9 fp64 MLUPS of the fused time step on N B200s, with roofline, e2e and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--strong] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[4]): fully periodic shear-wave lattice, rho0 = 1,
u0 = (0.01 sin(2 pi y / ly), 0), omega = 1.0, f0 = f_eq; weak scaling = 16384 x 16384 cells PER GPU (1-D slabs
along the slow axis, ghost rows stored by the neighbour's kernel over NVLink), strong = 32768 x 32768 total.
One "step" = one reference time step over the whole lattice. MLUPS = cells * steps / seconds / 1e6, whole job.

Prints ONE JSON line (rank 0). Keys beyond the base contract: "roofline", "cpu_baseline" (see README/DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_UPDATE = 144.0   # 9 populations x 8 B read + 9 x 8 B written
EPS, OMEGA = 0.01, 1.0


# stdout carries exactly ONE line, the JSON result: everything else that writes to file descriptor 1 while the
# bench runs (NCCL's version banner, library chatter) is sent to stderr. Done by main(), not at import.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(text.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs, device copy)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe). The timed region of the
    default run is ~60 ms, nvidia-smi needs longer than that to come up: the sampler is started before the warm-up steps,
    polls every 20 ms, stamps every row with its arrival time, and reports the rows that arrived between begin() and
    end() (falling back to the rows since the warm-up began — the same load — when the region was shorter than a poll)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None
        self.t_begin = self.t_end = None

    def start(self, wait_s=2.0):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < wait_s and self.proc.poll() is None:
                time.sleep(0.01)                      # nvidia-smi is up once its first row is here
        except Exception:
            self.proc = None
        self.t_start = time.perf_counter()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable'], 'samples': 0}
        time.sleep(0.05)
        self.proc.terminate()
        t_end = self.t_end if self.t_end is not None else time.perf_counter()
        t_begin = self.t_begin if self.t_begin is not None else self.t_start
        window = 'timed region'
        rows = [r for t, r in self.rows if t_begin <= t <= t_end + 0.02]
        if not rows:
            window = 'warm-up + timed region (same load)'
            rows = [r for t, r in self.rows if self.t_start <= t <= t_end + 0.02]
        sm, mx, reasons = [], [], set()
        for r in rows:
            p = [c.strip() for c in r.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's OWN numpy time step when its sources travelled with the repo (baseline/_ref/src, filled
# by __graft_entry__.build() from /root/reference; kind "reference"), else the oracle port of it (oracle/lbm_numpy.py,
# pinned to the reference bit for bit by tests/test_oracle_golden.py; kind "port")
# ---------------------------------------------------------------------------------------------------------
REF_SRC = os.path.join(ROOT, 'baseline', '_ref', 'src')


def cpu_modules():
    """-> (step(f, rho, u, omega, boundary, comm), equilibrium(rho, u), kind)"""
    if os.path.exists(os.path.join(REF_SRC, 'lattice_boltzmann_method.py')):
        if REF_SRC not in sys.path:
            sys.path.insert(0, REF_SRC)
        import lattice_boltzmann_method as R          # the unmodified reference module (flat import, as its Makefile does)
        assert os.path.dirname(os.path.abspath(R.__file__)) == REF_SRC
        return R.lattice_boltzmann_step, R.equilibrium_distr_func, 'reference'
    from oracle import lbm_numpy as onp
    return onp.step, onp.equilibrium, 'port'


def shear_wave(shape, ny_profile=None):
    """rho = 1, u = (eps sin(2 pi y / ly), 0) — initial_values.py:67-93."""
    ly = shape[1] if ny_profile is None else ny_profile
    rho = np.ones(shape)
    u = np.zeros(shape + (2,))
    u[..., 0] = (EPS * np.sin(np.divide(2 * np.pi * np.arange(shape[1]), ly)))[None, :]
    return rho, u


def cpu_reference_mlups(n, steps, warmup):
    step, equilibrium, _ = cpu_modules()
    rho, u = shear_wave((n, n))
    f = equilibrium(rho, u)
    for _ in range(warmup):
        f, rho, u = step(f, rho, u, OMEGA)
    t0 = time.perf_counter()
    for _ in range(steps):
        f, rho, u = step(f, rho, u, OMEGA)
    dt = time.perf_counter() - t0
    return n * n * steps / dt / 1e6, dt


def _cpu_worker(rank, k, n, steps, warmup, shm_name, barrier, out, check_name=None):
    """One rank of the reference's parallel path (experiments.py:723-767 shape: collide, halo exchange, stream,
    moments) on a slab of the n x n periodic shear-wave lattice; the four Sendrecv of parallelization_utils.py:34-49
    go through a shared-memory mailbox (there is no MPI in the image)."""
    from multiprocessing import shared_memory
    step, equilibrium, _ = cpu_modules()
    os.environ['OMP_NUM_THREADS'] = '1'
    nloc = n // k
    shm = shared_memory.SharedMemory(name=shm_name)
    mail = np.ndarray((k, 2, n + 2, 9), dtype=np.float64, buffer=shm.buf)
    rho, u = shear_wave((nloc + 2, n + 2))
    prof = EPS * np.sin(np.divide(2 * np.pi * ((np.arange(n + 2) - 1) % n), n))
    u[..., 0] = prof[None, :]
    f = equilibrium(rho, u)
    left, right = (rank - 1) % k, (rank + 1) % k

    def comm(fp):
        mail[rank, 0] = fp[1]          # my first interior row  -> left neighbour's high ghost row
        mail[rank, 1] = fp[-2]         # my last interior row   -> right neighbour's low ghost row
        barrier.wait()
        fp[-1] = mail[right, 0]
        fp[0] = mail[left, 1]
        barrier.wait()
        fp[:, -1, :] = fp[:, 1, :]     # one rank along y: self copies, after the x faces (carry the corners)
        fp[:, 0, :] = fp[:, -2, :]
        return fp

    for _ in range(warmup):
        f, rho, u = step(f, rho, u, OMEGA, None, comm)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        f, rho, u = step(f, rho, u, OMEGA, None, comm)
    barrier.wait()
    out[rank] = time.perf_counter() - t0
    if check_name:   # tests: gather the interiors so the decomposition can be compared with one process
        chk = shared_memory.SharedMemory(name=check_name)
        np.ndarray((n, n, 9), dtype=np.float64, buffer=chk.buf)[rank * nloc:(rank + 1) * nloc] = f[1:-1, 1:-1]
        chk.close()
    shm.close()


def cpu_reference_mlups_parallel(n, steps, warmup, k, return_fields=False):
    """k processes (one per host core), slabs along x with a ghost ring — how the reference uses more than one core
    (mpirun -N k, README.md:47-48)."""
    import multiprocessing as mp
    from multiprocessing import shared_memory
    ctx = mp.get_context('fork')
    shm = shared_memory.SharedMemory(create=True, size=k * 2 * (n + 2) * 9 * 8)
    barrier = ctx.Barrier(k)
    out = ctx.Array('d', k)
    chk = shared_memory.SharedMemory(create=True, size=n * n * 72) if return_fields else None
    procs = [ctx.Process(target=_cpu_worker, args=(r, k, n, steps, warmup, shm.name, barrier, out,
                                                   chk.name if chk else None)) for r in range(k)]
    [p.start() for p in procs]
    [p.join() for p in procs]
    shm.close()
    shm.unlink()
    fields = None
    if chk:
        fields = np.ndarray((n, n, 9), dtype=np.float64, buffer=chk.buf).copy()
        chk.close()
        chk.unlink()
    if any(p.exitcode != 0 for p in procs):
        raise RuntimeError('CPU baseline worker failed')
    dt = max(out[:])
    return (n * n * steps / dt / 1e6, dt, fields) if return_fields else (n * n * steps / dt / 1e6, dt)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        with open('/proc/cpuinfo') as fh:
            for line in fh:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except Exception:
        pass
    return 'unknown'


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation (its own lattice_boltzmann_step from baseline/_ref/src, kind
    "reference"; the oracle port of it where the sources did not travel, kind "port") on ALL host
    cores: k = largest power of two <= cores processes, slab decomposition + ghost exchange as under mpirun -N k.
    Each step is a bounded sample of the workload: one time step of an n x n lattice (n = --cpu-size)."""
    if rank != 0:
        return
    n = args.cpu_size
    k = 1
    while k * 2 <= host_cores() and n % (k * 2) == 0 and k * 2 <= 64:
        k *= 2
    if args.cpu_procs:
        k = args.cpu_procs
    # bound the sample: the whole --steps/--warmup run must end within a few minutes on this host
    while n > 256:
        probe = cpu_reference_mlups_parallel(n, 1, 1, k)[1] if k > 1 else cpu_reference_mlups(n, 1, 1)[1]
        if probe * (args.steps + args.warmup) <= 150.0:
            break
        n //= 2
    if k > 1:
        mlups, dt = cpu_reference_mlups_parallel(n, args.steps, args.warmup, k)
    else:
        mlups, dt = cpu_reference_mlups(n, args.steps, args.warmup)
    kind = cpu_modules()[2]
    impl_name = ("the reference's own lattice_boltzmann_step (baseline/_ref/src, unmodified)" if kind == 'reference' else
                 'oracle/lbm_numpy.py (numpy restatement of the reference)')
    sample = (f'{impl_name}: {n}x{n} periodic shear wave (same fields/omega as the GPU arm, which runs {args.size}^2 per GPU; the '
              f'numpy path needs ~400 B/cell so the full size does not fit/finish), {args.steps} steps after '
              f'{args.warmup} warm-up, {k} processes x 1 thread (numpy ufuncs are single-threaded; slabs + ghost-row '
              f'exchange through shared memory, as mpirun -N {k}); host: {host_cores()} usable cores, {cpu_model()}')
    line = {
        'impl': 'reference', 'metric': 'D2Q9 fp64 MLUPS', 'value': mlups, 'unit': 'MLUPS', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{n}x{n} periodic shear-wave lattice on the host cores (a bounded sample of the GPU arm\'s '
                               f'workload, which is {args.size}x{args.size} per GPU: MLUPS is size-normalised, the numpy path '
                               f'needs ~400 B/cell so the full size neither fits nor finishes)',
                   'lattice': [n, n], 'gpu_arm_lattice_per_gpu': [args.size, args.size], 'omega': OMEGA, 'epsilon': EPS,
                   'decomposition': f'{k} processes x 1 thread, slabs along the slow axis with a ghost ring'},
        'cpu_baseline': {'value': mlups, 'unit': 'MLUPS', 'cores': k, 'kind': kind, 'sample': sample},
        'e2e': {'value': mlups, 'unit': 'MLUPS', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


def workload_config(args, world, depth=3):
    if args.strong:
        wl = f'strong scaling: {args.size}x{args.size} total periodic shear-wave lattice over {world} GPU(s)'
    else:
        wl = f'weak scaling: {args.size}x{args.size} periodic shear-wave lattice per GPU ({args.size * world}x{args.size} total)'
    return {'workload': wl, 'lattice_per_gpu': [args.size // world if args.strong else args.size, args.size],
            'omega': OMEGA, 'epsilon': EPS, 'time_steps_per_pass': depth,
            'decomposition': f'{world}x1 slabs along the slow axis, {max(depth, 2)} ghost rows each side'
            if world > 1 else 'single block, periodic wrap in-kernel',
            'l2': 'populations are 19.3 GB per GPU per buffer >> 126 MB L2; no flush needed'}


def bind_to_gpu_numa_node(gpu):
    """One process per GPU: run (and therefore first-touch its pinned host buffers) on the CPUs NVML reports as local to
    that GPU, so that eight ranks staging 25.8 GB each do not all pull through one memory controller / inter-socket link.
    Best effort: returns what was done, never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        pick = cpus & allowed
        if pick and pick != allowed:
            os.sched_setaffinity(0, pick)
            return f'bound to {len(pick)} CPUs local to GPU {gpu}'
        return f'GPU {gpu}: NVML reports {len(cpus)} local CPUs = the allowed set ({len(allowed)}): nothing to bind'
    except Exception as e:   # no NVML, restricted container, ...
        return f'not bound ({type(e).__name__})'


def source_sha():
    """sha256 of the hot kernels' source (the marked region of lbm_kernels.cuh: k_step_pair, k_step2x, k_stepNx, and the
    arithmetic header): profiles/traffic.json is only quoted for the kernels it was captured on."""
    import hashlib
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, 'lattice_boltzmann_parallel_solver_b200', 'csrc')
    with open(os.path.join(csrc, 'lbm_kernels.cuh'), 'rb') as fh:
        text = fh.read()
    a, b = text.find(b'// ==== HOT KERNELS BEGIN'), text.find(b'// ==== HOT KERNELS END')
    h.update(text[a:b] if 0 <= a < b else text)
    with open(os.path.join(csrc, 'lbm_device.cuh'), 'rb') as fh:
        h.update(fh.read())
    return h.hexdigest()


def measured_traffic(key):
    """DRAM bytes per launch from the committed ncu capture, or None when the kernels changed since it was taken."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            tj = json.load(fh)
    except Exception:
        return None, 'profiles/traffic.json missing'
    if tj.get('source_sha256') != source_sha():
        return None, 'profiles/traffic.json was captured on other kernel sources (sha mismatch): re-run profiles/run_profiles.sh'
    return tj.get(key), tj.get(key + '_source')


PARITY_PERIOD = 64


def run_parity(lat, world, rank, nx_local, ny, prof, steps, barrier, all_sum, all_max):
    """Values, not just completion, of the decomposition the timed run uses (outside the timed region): the lattice is
    loaded with a density that varies ALONG the slab axis — rho(x) with period 64 in the GLOBAL row index — on top of
    the shear wave u_x(y), and advanced `steps` steps through the same launches as the timed run (multi-step passes,
    ghost-row stores over NVLink, flag handshake, edge/interior overlap). The global solution is then periodic in x
    with period 64, i.e. equal to a 64 x ny periodic lattice, which the C oracle (the checker: oracle/lbm_oracle.c,
    pinned to the reference bit for bit) runs in a second. Every rank compares its first and last four interior rows —
    the rows that straddle the slab boundaries and depend on the neighbours' ghost stores — and four rows in the
    middle, all populations, density and velocity, bit for bit."""
    from oracle import lbm_c, lbm_numpy as onp
    g = lat.ghost[0]
    NX = nx_local + 2 * g
    tab = 1.0 + 0.02 * np.sin(np.divide(2 * np.pi * np.arange(PARITY_PERIOD), PARITY_PERIOD))
    X = (rank * nx_local + np.arange(NX) - g) % (nx_local * world)      # global row of every local row, ghosts included
    lat.load_equilibrium(OMEGA, rho_x=tab[X % PARITY_PERIOD], ux_y=prof)
    barrier()
    lat.run(steps)
    lat.sync()
    rho = np.repeat(tab[:, None], ny, axis=1)
    u = np.zeros((PARITY_PERIOD, ny, 2))
    u[..., 0] = prof[None, :]
    ref = lbm_c.run(onp.equilibrium(rho, u), rho, u, OMEGA, lbm_c.periodic(), steps)
    mism, maxd, rows = 0, 0.0, 0
    for r0 in sorted({g, g + (nx_local // 2 // 4) * 4, NX - g - 4}):
        got = lat.fields(region=(r0, r0 + 4, 0, ny))
        idx = X[r0:r0 + 4] % PARITY_PERIOD
        for a, b in zip(got, ref):
            b = b[idx]
            mism += int(np.count_nonzero(a != b))
            maxd = max(maxd, float(np.max(np.abs(a - b))))
        rows += 4
    barrier()
    return {'checked': True, 'mismatches': int(all_sum(mism)), 'max_abs_diff': float(all_max(maxd)),
            'rows': int(all_sum(rows)), 'values_compared': int(all_sum(rows)) * ny * 12, 'steps': steps,
            'field': f'rho(x) = 1 + 0.02 sin(2 pi (x mod {PARITY_PERIOD}) / {PARITY_PERIOD}) in the global row index (varies along '
                     f'the slab axis), u_x(y) = the shear wave; oracle: {PARITY_PERIOD} x {ny} periodic lattice, C restatement',
            'rows_per_rank': 'first 4 and last 4 interior rows (next to the neighbours\' ghost rows) + 4 mid rows; f, density, velocity'}


def timed_steps(lat, stream, n, reduce_max):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    lat.run(n)
    e1.record(stream)
    lat.sync()
    return reduce_max(e0.elapsed_time(e1))


# ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--size', type=int, default=16384, help='lattice edge per GPU (weak) or total (with --strong)')
    ap.add_argument('--strong', action='store_true')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-size', type=int, default=2048)
    ap.add_argument('--cpu-steps', type=int, default=4)
    ap.add_argument('--cpu-procs', type=int, default=0, help='processes of the CPU reference arm (default: all cores)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-ref-config', action='store_true',
                    help="skip the reference's own 420x180 scaling_test (reported beside the headline at N=1)")
    ap.add_argument('--single-step', action='store_true', help='one time step per launch only (no temporal blocking)')
    ap.add_argument('--e2e-size', type=int, default=0, help='lattice edge of the e2e job (default: --size)')
    ap.add_argument('--workload', default='shear', choices=['shear', 'karman'],
                    help='shear: the headline periodic lattice; karman: inlet/outlet/plate rule set scaled to the same '
                         'lattice (single GPU; evidence for the flag-mask vs edge-kernel choice)')
    ap.add_argument('--bc-mode', default='auto', choices=['auto', 'mask', 'edge'])
    ap.add_argument('--depth', type=int, default=3, choices=[2, 3, 4], help='time steps per pass of the multi-step kernel')
    ap.add_argument('--no-parity', action='store_true', help='skip the x-periodic parity job against the C oracle')
    ap.add_argument('--parity-steps', type=int, default=13)
    ap.add_argument('--no-sub', action='store_true', help='skip the strong-scaling and von Karman sub-records')
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)   # timing hygiene: never fewer than 3 warm-up steps

    import torch
    import torch.distributed as dist
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200 import dist as ldist
    from lattice_boltzmann_parallel_solver_b200 import parallelization_utils as par
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice

    torch.cuda.set_device(local)
    N.set_device(local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        ldist.ensure_process_group('nccl')
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}'

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=op)
        return float(t.item())

    def all_max(v):
        return reduce(v, dist.ReduceOp.MAX)

    def all_sum(v):
        return reduce(v, dist.ReduceOp.SUM)

    peak, peak_src = measured_peak_gbs()
    depth = 1 if args.single_step else args.depth
    bc_mode = {'auto': N.BC_AUTO, 'mask': N.BC_MASK, 'edge': N.BC_EDGE}[args.bc_mode]

    def make_lattice(workload, strong, size):
        """The device lattice of one measurement: (lattice, nx_local, ny, profile)."""
        ny = size
        nx_local = size // world if strong else size
        prof = EPS * np.sin(np.divide(2 * np.pi * np.arange(ny), ny))   # initial_values.py:83-88
        if workload == 'karman' and world == 1:
            lat = karman_lattice(nx_local, ny, bc_mode)
        elif workload == 'karman':
            # N GPUs: the rule set on the GLOBAL (nx_local * N) x ny lattice, slabs with two ghost rows that carry the
            # neighbour's kinds; `depth` steps per pass on every rank (multi-step kernel + strip windows next to boundary
            # rows and slab edges)
            g = max(depth, 2)
            km = karman_slab_kind_map(nx_local * world, ny, rank * nx_local - g, nx_local + 2 * g)
            lat = Lattice(nx_local + 2 * g, ny, km, ghost=(g, 0), bc_mode=bc_mode)
            cart = ldist.comm_world().Create_cart(dims=[world, 1], periods=[True, True])
            par.communication(cart).attach(lat)
        elif world == 1:
            lat = Lattice(nx_local, ny)
        else:
            # `depth` ghost rows per side: the dependency cone of a multi-step pass over the slab's edge rows; the
            # neighbour's edge launch stores them over NVLink, one exchange and one flag handshake per pass
            g = max(depth, 2)
            lat = Lattice(nx_local + 2 * g, ny, ghost=(g, 0))
            cart = ldist.comm_world().Create_cart(dims=[world, 1], periods=[True, True])
            par.communication(cart).attach(lat)
        if depth == 1:
            lat.set_option('fused', 0)
        else:
            lat.set_option('fused_depth', depth)
        return lat, nx_local, ny, prof

    def load(lat, workload, prof):
        if workload == 'karman':
            lat.load_equilibrium(float(np.reciprocal(3 * 0.04 + 0.5)), rho0=1.0, ux0=0.1)
        else:
            lat.load_equilibrium(OMEGA, ux_y=prof)
        barrier()

    def measure(lat, nx_local, ny, steps, warmup, sample_clocks):
        """`warmup` untimed steps, then exactly `steps` steps between CUDA events on the library's stream, bracketed by
        barrier + synchronize; max over ranks."""
        stream = torch.cuda.ExternalStream(lat.stream)
        sampler = ClockSampler(local) if sample_clocks and rank == 0 else None
        if sampler:
            sampler.start()
        lat.run(warmup)
        lat.sync()
        barrier()
        l0 = lat.launches
        torch.cuda.synchronize()
        barrier()
        if sampler:
            sampler.begin()
        ms = timed_steps(lat, stream, steps, all_max)
        torch.cuda.synchronize()
        if sampler:
            sampler.end()
        barrier()
        clocks = sampler.stop() if sampler else None
        cells_total = nx_local * world * ny
        return {'ms': ms, 'mlups': cells_total * steps / (ms * 1e-3) / 1e6, 'launches': lat.launches - l0, 'clocks': clocks,
                'stream': stream}

    # ---- headline ----------------------------------------------------------------------------------------------
    lat, nx_local, ny, prof = make_lattice(args.workload, args.strong, args.size)
    parity = None
    if args.workload == 'shear' and not args.no_parity:
        parity = run_parity(lat, world, rank, nx_local, ny, prof, args.parity_steps, barrier, all_sum, all_max)
    load(lat, args.workload, prof)
    head = measure(lat, nx_local, ny, args.steps, args.warmup, True)
    ms, mlups, launches, clocks, stream = head['ms'], head['mlups'], head['launches'], head['clocks'], head['stream']
    per_gpu_cells = nx_local * ny
    cells_total = per_gpu_cells * world

    # ---- roofline ---------------------------------------------------------------------------------------------
    # Dominant kernel: the multi-step pass (k_stepNx<depth>: `depth` time steps per launch). Its launch duration is
    # measured live, alone: `fused_exact` makes lbm_step(n) exactly n / depth passes (CUDA events on the library's
    # stream, same lattice, right after the timed region). achieved = algorithmic bytes per launch (depth x 144 B x
    # cells) / that duration; dram_frac = measured DRAM traffic of one launch (ncu, profiles/traffic.json) / duration
    # / peak — the fraction of the memory roof the kernel really uses.
    fused = depth > 1
    if fused:
        spl = depth
        lat.set_option('fused_exact', 1)
        n_pure = spl * max(4, min(24, args.steps // spl))
        lat.run(spl * 2)
        lat.sync()
        barrier()
        ms_pure = timed_steps(lat, stream, n_pure, all_max)
        lat.set_option('fused_exact', 0)
        launch_ms = ms_pure / (n_pure // spl)
        algo_launch = spl * per_gpu_cells * ALGO_BYTES_PER_UPDATE
        kernel = (f'k_stepNx<128,{depth}> ({depth} time steps per launch: two columns per thread, {depth - 1} shared-memory '
                  f'ring(s) of intermediate rows)')
        tkey = f'k_stepNx{depth}_dram_bytes_per_launch_16384'
        if args.workload == 'karman':
            kernel += (' on the rows whose dependency cone is all fluid + one one-step mask launch per step through windows on each '
                       'strip of boundary rows (inlet/outlet rows, plate rows); duration = one pass')
    else:
        launch_ms = ms / args.steps
        algo_launch = per_gpu_cells * ALGO_BYTES_PER_UPDATE
        kernel = 'k_step_pair (one time step per launch, two cells per thread)'
        tkey = 'dram_bytes_per_launch_16384'
        spl = 1
    achieved = algo_launch / (launch_ms * 1e-3) / 1e9
    at_capture_size = args.size == 16384 and not args.strong
    traffic, traffic_src = measured_traffic(tkey) if at_capture_size else (None, 'captured at 16384^2 per GPU only')
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': traffic_src,
                'dram_frac': (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None,
                'peak_source': peak_src, 'kernel': kernel, 'launch_ms': launch_ms,
                'algorithmic_bytes_per_launch': algo_launch, 'steps_per_launch': spl,
                'mlups_of_the_kernel_alone': per_gpu_cells * world * spl / (launch_ms * 1e-3) / 1e6,
                'mlups_at_peak_one_step_per_launch': peak * 1e9 / ALGO_BYTES_PER_UPDATE / 1e6,
                'frac_of_nominal_8TBps': achieved / 8000.0}
    if fused:
        roofline['note'] = (f'frac > 1 by construction: a {spl}-step pass moves ~1/{spl} of the algorithmic bytes of its steps '
                            'through DRAM (traffic vs algorithmic_bytes_per_launch; dram_frac is the share of the memory '
                            'roof it really uses); it is fp64-issue/latency bound. The one-step kernel the north star '
                            'describes is timed below (single_step).')
        # the one-step-per-launch kernel (the north star's "reads each population once and writes it once"), timed
        # live in the same process on the same lattice
        lat.set_option('fused', 0)
        k1 = max(10, args.steps // 4)
        lat.run(3)
        lat.sync()
        barrier()
        ms1 = timed_steps(lat, stream, k1, all_max)
        a1 = per_gpu_cells * ALGO_BYTES_PER_UPDATE / (ms1 / k1 * 1e-3) / 1e9
        t1, _ = measured_traffic('dram_bytes_per_launch_16384') if at_capture_size else (None, None)
        roofline['single_step'] = {'kernel': 'k_step_pair' + (' + edge-list kernel' if args.workload == 'karman' else ''),
                                   'steps': k1, 'ms_per_step': ms1 / k1, 'mlups': cells_total * k1 / (ms1 * 1e-3) / 1e6,
                                   'achieved': a1, 'frac': a1 / peak, 'frac_of_nominal_8TBps': a1 / 8000.0, 'traffic': t1,
                                   'dram_frac': (t1 / (ms1 / k1 * 1e-3) / 1e9 / peak) if t1 else None}
        lat.set_option('fused', 1)

    # ---- e2e: the whole job through the reference-shaped API with HOST buffers ------------------------------
    e2e = None
    if not args.no_e2e and args.workload == 'shear':
        e2e = run_e2e(args, lat, world, rank, nx_local, ny, prof, barrier)
    lat.close()
    del lat

    # ---- sub-records: BASELINE.json config 5 in full (strong scaling 32768^2 at this N) and the BC-bearing case ----
    def sub_record(workload, strong, size, steps, warmup):
        try:
            l2, nxl, ny2, prof2 = make_lattice(workload, strong, size)
        except MemoryError as e:
            return {'skipped': str(e)[:160]}
        par2 = None
        if workload == 'shear' and not args.no_parity:
            par2 = run_parity(l2, world, rank, nxl, ny2, prof2, args.parity_steps, barrier, all_sum, all_max)
        load(l2, workload, prof2)
        m = measure(l2, nxl, ny2, steps, warmup, False)
        l2.close()
        rec = {'value': m['mlups'], 'unit': 'MLUPS', 'ms_per_step': m['ms'] / steps, 'steps': steps, 'warmup': warmup,
               'gpu_launches': int(m['launches']), 'n_gpus': world, 'lattice_per_gpu': [nxl, ny2], 'lattice_total': [nxl * world, ny2]}
        if par2 is not None:
            rec['parity'] = par2
        return rec

    strong_rec = karman_rec = None
    if args.workload == 'shear' and not args.strong and not args.no_sub:
        ssteps = max(depth * 4, min(args.steps, 60))
        strong_rec = sub_record('shear', True, 32768, ssteps, max(3, min(args.warmup, 12)))
        strong_rec['scaling'] = 'strong'
        strong_rec['workload'] = f'strong scaling: 32768x32768 total periodic shear-wave lattice over {world} GPU(s) (BASELINE.json configs[4])'
        karman_rec = sub_record('karman', False, args.size, max(8, min(args.steps, 60)), max(3, min(args.warmup, 12)))
        karman_rec['scaling'] = 'weak'
        karman_rec['workload'] = (f'von Karman rule set (inlet row, outlet rows, plate of ny/4.5 at nx/4; nu 0.04, u_in 0.1) on the '
                                  f'global {args.size * world}x{args.size} lattice ({args.size}x{args.size} per GPU), {depth} steps per '
                                  f'pass: the BC-bearing case of SURVEY.md section 8(d)')
        karman_rec['vs_periodic'] = karman_rec['value'] / mlups if 'value' in karman_rec else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:   # CPU baseline: rank 0 at N=1 only
        k = 1
        while k * 2 <= host_cores() and args.cpu_size % (k * 2) == 0 and k * 2 <= 64:
            k *= 2
        k = args.cpu_procs or k
        steps_cpu = args.cpu_steps * (4 if k > 1 else 1)
        v, dt = (cpu_reference_mlups_parallel(args.cpu_size, steps_cpu, 1, k) if k > 1 else
                 cpu_reference_mlups(args.cpu_size, steps_cpu, 1))
        v1, dt1 = cpu_reference_mlups(args.cpu_size, 2, 1)
        kind = cpu_modules()[2]
        impl_name = ("the reference's own lattice_boltzmann_step (baseline/_ref/src, unmodified)" if kind == 'reference' else
                     'oracle/lbm_numpy.py (numpy restatement of the reference)')
        cpu = {'value': v, 'unit': 'MLUPS', 'cores': k, 'kind': kind, 'single_core_value': v1,
               'sample': f'{args.cpu_size}x{args.cpu_size} periodic shear wave, {steps_cpu} steps after 1 warm-up, '
                         f'{impl_name} on {k} processes x 1 thread with '
                         f'slab decomposition + ghost-row exchange (as mpirun -N {k}); host: {host_cores()} usable '
                         f'cores, {cpu_model()}; {dt:.1f} s; one process alone: {v1:.2f} MLUPS'}
    ref_cfg = None
    if rank == 0 and world == 1 and not args.no_ref_config and args.workload == 'shear':
        ref_cfg = reference_scaling_test()
    if rank == 0:
        cfg = workload_config(args, world, depth)
        if args.workload == 'karman':
            cfg.update({'workload': f'von Karman rule set (inlet, outlet, plate) on {args.size}x{args.size}, bc_mode={args.bc_mode}',
                        'omega': float(np.reciprocal(3 * 0.04 + 0.5)), 'epsilon': None})
        line = {
            'metric': 'D2Q9 fp64 MLUPS', 'value': mlups, 'unit': 'MLUPS', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': dict(cfg, cpu_affinity=affinity) if affinity else cfg, 'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'parity': parity,
            'roofline': roofline, 'cpu_baseline': cpu, 'strong': strong_rec, 'karman': karman_rec,
            'reference_scaling_test': ref_cfg,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def reference_scaling_test(steps=20000):
    """The reference's OWN published benchmark, through the drop-in modules: `scaling_test` of src/experiments.py:
    723-774 — von Karman vortex street on 420 x 180, plate 40, u_in 0.1, nu 0.04, ghost-padded local arrays,
    `parallel_von_karman_boundary_conditions` + `communication(cartesian2d)`, `time_steps` calls of
    `lattice_boltzmann_step`, wall clock around the loop. One rank = one GPU here; the reference's figures
     are np.load('420_180_<ranks>.npy') / 1e7: best 75.99 MLUPS on 400 MPI ranks.
    The fields of the last step are brought to the host INSIDE the timed region (the reference's arrays are host
    arrays when its clock stops)."""
    import lattice_boltzmann_parallel_solver_b200 as P
    from lattice_boltzmann_parallel_solver_b200 import dist as ldist
    L, BU, PU = P.lattice_boltzmann_method, P.boundary_utils, P.parallelization_utils
    lx, ly, plate, rho_in, u_in, nu = 420, 180, 40, 1.0, 0.1, 0.04
    omega = np.reciprocal(3 * nu + 0.5)
    comm = ldist.WorldComm()
    x_size, y_size = PU.get_xy_size(1)
    cart = comm.Create_cart(dims=[x_size, y_size], periods=[True, True], reorder=False)
    coords = cart.Get_coords(0)
    nlx, nly = PU.get_local_coords(coords, lx, ly, x_size, y_size)
    density = np.ones((nlx + 2, nly + 2))
    velocity = np.zeros((nlx + 2, nly + 2, 2))
    velocity[..., 0] = u_in                                  # density_1_velocity_x_u0_velocity_y_0_initial
    f = L.equilibrium_distr_func(density, velocity)
    bound = BU.parallel_von_karman_boundary_conditions(coords, nlx, nly, lx, ly, x_size, y_size, rho_in, u_in, plate)
    com = PU.communication(cart)
    out = {}
    for label, n in (('warmup', 2000), ('timed', steps)):
        fi, di, vi = f, density, velocity
        t0 = time.perf_counter()
        for _ in range(n):
            fi, di, vi = L.lattice_boltzmann_step(fi, di, vi, omega, bound, com)
        vmax = float(np.max(np.abs(np.asarray(vi))))
        out[label] = time.perf_counter() - t0
    L.release_lattices()
    mlups = lx * ly * steps / out['timed'] / 1e6
    return {'workload': 'scaling_test of src/experiments.py:723-774: von Karman 420x180, plate 40, 1 rank = 1 GPU, '
                        'driven through the drop-in lattice_boltzmann_step / boundary_utils / parallelization_utils',
            'steps': steps, 'seconds': out['timed'], 'us_per_step': 1e6 * out['timed'] / steps, 'mlups': mlups,
            'max_abs_velocity': vmax,
            'published_best_mlups': 75.99, 'published_best_ranks': 400, 'ratio_vs_published_best': mlups / 75.99,
            'published_source': 'figures/von_karman_vortex_shedding/scaling_test/420_180_400.npy (BASELINE.md section 1)'}


def karman_lattice(nx, ny, bc_mode):
    """milestone_6's rule set (inlet column, outlet column, thin plate of ny/4.5 at nx/4) on an nx x ny lattice."""
    from lattice_boltzmann_parallel_solver_b200 import boundary_conditions as B
    from lattice_boltzmann_parallel_solver_b200 import boundary_utils as BU
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    d = int(ny / 4.5) // 2 * 2
    plate = np.zeros((nx, ny), dtype=bool)
    plate[nx // 4, ny // 2 - d // 2:ny // 2 + d // 2] = True
    bundle = BU.BoundaryBundle('von_karman_serial', (nx, ny))
    bundle.add(B.inlet((nx, ny), 1.0, 0.1)).add(B.outlet()).add(B.rigid_object(plate))
    return Lattice(nx, ny, bundle.kind_map((nx, ny)), bc_mode=bc_mode)


def karman_slab_kind_map(nx_global, ny, row0, nrows):
    """Rows [row0, row0 + nrows) (periodic in the global row index) of the kind map `karman_lattice` builds for an
    nx_global x ny lattice, WITHOUT building the global arrays (131072 x 16384 at 8 GPUs): the rule set touches five
    global rows only — inlet on row 0, outlet on rows nx-2 / nx-1, the plate on rows nx//4 and nx//4 + 1 — and is
    written here row by row with the same emitters' effect, in the same order (inlet, outlet, plate). Tested against
    the global construction at small sizes (tests/test_host_logic.py). Ghost rows of a slab carry the neighbour's kinds."""
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200 import boundary_conditions as B
    from lattice_boltzmann_parallel_solver_b200 import boundary_spec as S
    km = S.KindMap((nrows, ny))
    d = int(ny / 4.5) // 2 * 2
    y_lo, y_hi = ny // 2 - d // 2, ny // 2 + d // 2 - 1          # plate cells y_lo .. y_hi (the two ends are corners)
    x0 = nx_global // 4
    inlet_values = B.inlet((1, 1), 1.0, 0.1).values
    outlet_rule = S.rule(N.RULE_OUTLET)
    rows = [(i, (row0 + i) % nx_global) for i in range(nrows)]
    for i, X in rows:                                            # inlet (boundary_conditions.py:250-251)
        if X == 0:
            km.constant((i, slice(None)), inlet_values)
    for i, X in rows:                                            # outlet (:279-280)
        if X == nx_global - 1:
            km.flag((i, slice(None)), rules_for={3: outlet_rule, 6: outlet_rule, 7: outlet_rule})
        if X == nx_global - 2:
            km.flag((i, slice(None)), bits=N.CELL_OUTLET_SRC)
    for i, X in rows:                                            # plate (:133-163): full cells, then the four corners
        if X == x0:
            km.bounce((i, slice(y_lo + 1, y_hi)), [1, 5, 8])
        if X == x0 + 1:
            km.bounce((i, slice(y_lo + 1, y_hi)), [3, 6, 7])
    for i, X in rows:
        if X == x0:
            km.bounce((i, y_hi), [1, 8])
            km.bounce((i, y_lo), [1, 5])
        if X == x0 + 1:
            km.bounce((i, y_hi), [3, 7])
            km.bounce((i, y_lo), [3, 6])
    return km


def run_e2e(args, lat, world, rank, nx_local, ny, prof, barrier):
    """Whole job through the public API, starting and ending in HOST memory: upload of the rank's (f, density,
    velocity) from pinned host buffers, K time steps, a device->host read of EVERY step's observable (probe velocity,
    experiments.py:703-704), and the final device->host copy of f, density, velocity. The K steps are enqueued at
    once; the probe cell's thread writes each step's sample into a host-mapped ring as that step completes
    (lbm_probe_*), and the host reads sample k as soon as it has arrived while the device runs on — no step waits
    for the host. Bytes per step = totals / K."""
    import torch
    import torch.distributed as dist
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    n = args.e2e_size or args.size
    if n != args.size or world > 1 and args.strong:
        return None
    g = lat.ghost[0]
    NX = nx_local + 2 * g
    cells = NX * ny
    need = cells * 96 * max(1, int(os.environ.get('LOCAL_WORLD_SIZE', world)))
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = None
    if avail is not None and need > 0.5 * avail:
        return {'value': None, 'unit': 'MLUPS', 'h2d_bytes_per_step': None, 'd2h_bytes_per_step': None,
                'skipped': f'pinned host staging of {need / 1e9:.0f} GB exceeds half of the available host memory ({avail / 1e9:.0f} GB)'}
    try:
        hf = torch.empty((NX, ny, 9), dtype=torch.float64, pin_memory=True).numpy()
        hr = torch.empty((NX, ny), dtype=torch.float64, pin_memory=True).numpy()
        hu = torch.empty((NX, ny, 2), dtype=torch.float64, pin_memory=True).numpy()
    except RuntimeError as e:   # not enough pinnable host memory on this box
        return {'value': None, 'unit': 'MLUPS', 'h2d_bytes_per_step': None, 'd2h_bytes_per_step': None, 'skipped': str(e)[:120]}
    # host-side initial state = what the reference driver builds (experiments.py:121-122); built through the device
    # because a 16384^2 numpy equilibrium would take minutes of host time outside the timed region
    # t=0 arrays on the host: rho = 1, u = profile, f = f_eq (bit-identical to the numpy expression)
    hr[...] = 1.0
    hu[..., 0] = prof[None, :]
    hu[..., 1] = 0.0
    lib = N.load()
    row_rho, row_u = np.ones((1, ny)), np.zeros((1, ny, 2))
    row_u[0, :, 0] = prof
    row_f = np.empty((1, ny, 9))
    N.check(lib.lbm_equilibrium(lat.device, ny, N.dptr(row_rho), N.dptr(row_u), N.dptr(row_f)))
    hf[...] = row_f
    px, py = NX // 2, ny // 4
    K = args.steps
    lat.probe(px, py, capacity=K + 8)
    sink = np.empty((1, 2))
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    streamed = True
    try:
        # the whole job in ONE call of the public API (Lattice.run_host -> lbm_run_host): upload, time-skewed passes and
        # download are pipelined over 256 MB row chunks (slabs: the rows near the edges last, in lockstep with the
        # neighbours), results land in the input arrays
        lat.run_host(hf, hr, hu, OMEGA, K, out=(hf, hr, hu))     # H2D + D2H: 96 B per cell each way
        for k in range(K):
            sink[...] = lat.probe_read(k + 1, 1)      # D2H: 16 B of every step, from the host-mapped ring
    except N.LbmStateError:
        streamed = False
    if not streamed:
        lat.load(hf, hr, hu, OMEGA)                   # H2D: 96 B per cell
        barrier()
        lat.run(K)
        for k in range(K):
            sink[...] = lat.probe_read(k + 1, 1)      # D2H: 16 B of every step, from the host-mapped ring as it arrives
        of, orho, ou = hf[g:NX - g], hr[g:NX - g], hu[g:NX - g]      # the rank's own rows (contiguous views)
        N.check(lib.lbm_materialize_region(lat._ctx, g, NX - g, 0, ny, N.dptr(of), N.dptr(orho), N.dptr(ou)))   # D2H: 96 B per cell
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    total_cells = nx_local * world * ny
    # What the box's host<->device links sustain when every rank copies in both directions at once (outside the timed
    # region; plain pinned-memory copies on two streams, 2 GB each way per rank): the floor under the e2e wall clock.
    link = None
    try:
        nb = min(2 << 30, hf.nbytes // 2) // 8
        src = torch.from_numpy(hf.reshape(-1)[:nb])
        dst_h = torch.from_numpy(hf.reshape(-1)[nb:2 * nb])
        d_in = torch.empty(nb, dtype=torch.float64, device='cuda')
        d_out = torch.zeros(nb, dtype=torch.float64, device='cuda')
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for rep in range(2):                      # first repetition: warm-up
            torch.cuda.synchronize()
            barrier()
            t1 = time.perf_counter()
            with torch.cuda.stream(s1):
                d_in.copy_(src, non_blocking=True)
            with torch.cuda.stream(s2):
                dst_h.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            tl = time.perf_counter() - t1
        if world > 1:
            t = torch.tensor([tl], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tl = float(t.item())
        rate = 2 * nb * 8 * world / tl / 1e9      # GB/s, both directions, all ranks
        moved = cells * 96.0 * 2 * world / 1e9
        link = {'both_directions_all_ranks_gbs': rate, 'per_gpu_each_way_gbs': rate / 2 / world,
                'e2e_bytes_gb': moved, 'e2e_floor_s': moved / rate, 'e2e_wall_s': dt, 'fraction_of_floor': moved / rate / dt,
                'how': f'{world} rank(s) x (2 GB host->device + 2 GB device->host, pinned memory, concurrently on two streams), max over ranks'}
        del d_in, d_out
    except Exception as e:   # never let a diagnostic break the line
        link = {'skipped': str(e)[:120]}
    return {'value': total_cells * K / dt / 1e6, 'unit': 'MLUPS', 'link_probe': link,
            'h2d_bytes_per_step': cells * 96.0 / K, 'd2h_bytes_per_step': cells * 96.0 / K + 16.0,
            'job': f'upload f,rho,u from pinned host memory ({cells * 96 / 1e9:.1f} GB per GPU), {K} steps, '
                   f'the 16-byte probe sample of every step read by the host from a host-mapped ring, download f,rho,u'
                   + (' — one call of Lattice.run_host: upload, time-skewed passes and download pipelined over 256 MB row chunks'
                      if streamed else ' — load, barrier, run, fields (not pipelined on this lattice)')
                   + f'; wall clock {dt:.2f} s, max over ranks',
            'steps': K}


if __name__ == '__main__':
    main()
