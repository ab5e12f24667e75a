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
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
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
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's numpy algorithm (the reference itself is Python and does not
# travel to the GPU box; oracle/lbm_numpy.py is pinned to it bit-for-bit by tests/test_oracle_golden.py)
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_mlups(n, steps, warmup):
    from oracle import lbm_numpy as onp
    rho, u = onp.sinusoidal_velocity_x((n, n), EPS)
    f = onp.equilibrium(rho, u)
    for _ in range(warmup):
        f, rho, u = onp.step(f, rho, u, OMEGA)
    t0 = time.perf_counter()
    for _ in range(steps):
        f, rho, u = onp.step(f, rho, u, OMEGA)
    dt = time.perf_counter() - t0
    return n * n * steps / dt / 1e6, dt


def _cpu_worker(rank, k, n, steps, warmup, shm_name, barrier, out, check_name=None):
    """One rank of the reference's parallel path (experiments.py:723-767 shape: collide, halo exchange, stream,
    moments) on a slab of the n x n periodic shear-wave lattice; the four Sendrecv of parallelization_utils.py:34-49
    go through a shared-memory mailbox (there is no MPI in the image)."""
    from multiprocessing import shared_memory
    from oracle import lbm_numpy as onp
    os.environ['OMP_NUM_THREADS'] = '1'
    nloc = n // k
    shm = shared_memory.SharedMemory(name=shm_name)
    mail = np.ndarray((k, 2, n + 2, 9), dtype=np.float64, buffer=shm.buf)
    rho, u = onp.sinusoidal_velocity_x((nloc + 2, n + 2), EPS)
    prof = EPS * np.sin(np.divide(2 * np.pi * ((np.arange(n + 2) - 1) % n), n))
    u[..., 0] = prof[None, :]
    f = onp.equilibrium(rho, u)
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
        f, rho, u = onp.step(f, rho, u, OMEGA, None, comm)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        f, rho, u = onp.step(f, rho, u, OMEGA, None, comm)
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
    """--impl reference: the reference's CPU algorithm (oracle port of its numpy path, kind "port") on ALL host
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
    sample = (f'{n}x{n} periodic shear wave (same fields/omega as the GPU arm, which runs {args.size}^2 per GPU; the '
              f'numpy path needs ~400 B/cell so the full size does not fit/finish), {args.steps} steps after '
              f'{args.warmup} warm-up, {k} processes x 1 thread (numpy ufuncs are single-threaded; slabs + ghost-row '
              f'exchange through shared memory, as mpirun -N {k}); host: {host_cores()} usable cores, {cpu_model()}')
    line = {
        'impl': 'reference', 'metric': 'D2Q9 fp64 MLUPS', 'value': mlups, 'unit': 'MLUPS', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, 1),
        'cpu_baseline': {'value': mlups, 'unit': 'MLUPS', 'cores': k, 'kind': 'port', 'sample': sample},
        'e2e': {'value': mlups, 'unit': 'MLUPS', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


def workload_config(args, world):
    if args.strong:
        wl = f'strong scaling: {args.size}x{args.size} total periodic shear-wave lattice over {world} GPU(s)'
    else:
        wl = f'weak scaling: {args.size}x{args.size} periodic shear-wave lattice per GPU ({args.size * world}x{args.size} total)'
    return {'workload': wl, 'lattice_per_gpu': [args.size // world if args.strong else args.size, args.size],
            'omega': OMEGA, 'epsilon': EPS, 'decomposition': f'{world}x1 slabs along the slow axis, 2 ghost rows each side'
            if world > 1 else 'single block, periodic wrap in-kernel',
            'l2': 'populations are 19.3 GB per GPU per buffer >> 126 MB L2; no flush needed'}


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
    if world > 1:
        ldist.ensure_process_group('nccl')
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}'

    ny = args.size
    nx_local = args.size // world if args.strong else args.size
    nx_global = nx_local * world
    prof = EPS * np.sin(np.divide(2 * np.pi * np.arange(ny), ny))   # initial_values.py:83-88

    def barrier():
        if world > 1:
            dist.barrier()

    if args.workload == 'karman':
        assert world == 1, 'the BC-bearing comparison is a single-GPU measurement'
        lat = karman_lattice(nx_local, ny, {'auto': N.BC_AUTO, 'mask': N.BC_MASK, 'edge': N.BC_EDGE}[args.bc_mode])
        args.no_e2e = True
    elif world == 1:
        lat = Lattice(nx_local, ny)
    else:
        # two ghost rows per side: the two-steps-per-pass kernel needs the depth-2 dependency cone of its edge rows
        lat = Lattice(nx_local + 4, ny, ghost=(2, 0))
        cart = ldist.comm_world().Create_cart(dims=[world, 1], periods=[True, True])
        par.communication(cart).attach(lat)
    if args.workload == 'karman':
        lat.load_equilibrium(float(np.reciprocal(3 * 0.04 + 0.5)), rho0=1.0, ux0=0.1)
    else:
        lat.load_equilibrium(OMEGA, ux_y=prof)
    if args.single_step:
        lat.set_option('fused', 0)
    barrier()

    stream = torch.cuda.ExternalStream(lat.stream)
    lat.run(args.warmup)
    lat.sync()
    barrier()
    l0 = lat.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    e0.record(stream)
    lat.run(args.steps)
    e1.record(stream)
    lat.sync()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lat.launches - l0
    if world > 1:
        t = torch.tensor([ms], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cells_total = nx_global * ny
    mlups = cells_total * args.steps / (ms * 1e-3) / 1e6

    # ---- roofline ---------------------------------------------------------------------------------------------
    # Dominant kernel of the timed region: k_step2x, which advances TWO time steps per launch (temporal blocking).
    # achieved = algorithmic bytes per launch (2 steps x 144 B x cells) / mean launch time (CUDA events above; the
    # region is `pairs` two-step launches + 1-2 one-step launches, all back to back on one stream).
    peak, peak_src = measured_peak_gbs()
    per_gpu_cells = nx_local * ny
    fused = not args.single_step
    step_ms = ms / args.steps
    if fused:
        algo_launch = 2 * per_gpu_cells * ALGO_BYTES_PER_UPDATE
        achieved = algo_launch / (2 * step_ms * 1e-3) / 1e9
        kernel = ('k_step2x<128> (two time steps per launch: two columns per thread, shared-memory ring of the '
                  'intermediate rows)')
        if args.workload == 'karman':
            kernel += (' on the rows whose two-step dependency cone is all fluid + two one-step mask launches through '
                       'a window on each strip of boundary rows (inlet/outlet rows, plate rows)')
    else:
        algo_launch = per_gpu_cells * ALGO_BYTES_PER_UPDATE
        achieved = algo_launch / (step_ms * 1e-3) / 1e9
        kernel = 'k_step_pair (one time step per launch, two cells per thread)'
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': None, 'peak_source': peak_src, 'kernel': kernel,
                'algorithmic_bytes_per_launch': algo_launch, 'steps_per_launch': 2 if fused else 1,
                'mlups_at_peak_one_step_per_launch': peak * 1e9 / ALGO_BYTES_PER_UPDATE / 1e6,
                'frac_of_nominal_8TBps': achieved / 8000.0}
    prof_path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(prof_path):
        try:
            with open(prof_path) as fh:
                tj = json.load(fh)
            roofline['traffic'] = tj.get('k_step2x_dram_bytes_per_launch_16384' if fused else 'dram_bytes_per_launch_16384')
            if fused:
                roofline['note'] = ('frac > 1 by construction: the kernel moves ~half the algorithmic bytes of its two '
                                    'steps through DRAM (traffic vs algorithmic_bytes_per_launch); it is issue/latency '
                                    'bound, not bandwidth bound. The one-step kernel the north star describes is timed '
                                    'below (single_step).')
        except Exception:
            pass
    # the one-step-per-launch kernel (the north star's "reads each population once and writes it once"), timed
    # live in the same process for the same lattice
    if fused:
        lat.set_option('fused', 0)
        k1 = max(10, args.steps // 4)
        lat.run(3)
        lat.sync()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        lat.run(k1)
        s1.record(stream)
        lat.sync()
        ms1 = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ms1], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms1 = float(t.item())
        a1 = per_gpu_cells * ALGO_BYTES_PER_UPDATE / (ms1 / k1 * 1e-3) / 1e9
        roofline['single_step'] = {'kernel': 'k_step_pair' + (' + edge-list kernel' if args.workload == 'karman' else ''), 'steps': k1, 'ms_per_step': ms1 / k1,
                                   'mlups': cells_total * k1 / (ms1 * 1e-3) / 1e6, 'achieved': a1, 'frac': a1 / peak,
                                   'frac_of_nominal_8TBps': a1 / 8000.0,
                                   'traffic': tj.get('dram_bytes_per_launch_16384') if os.path.exists(prof_path) else None}
        lat.set_option('fused', 1)

    # ---- e2e: the whole job through the reference-shaped API with HOST buffers ------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, lat, world, rank, nx_local, ny, prof, barrier)
    lat.close()

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
        cpu = {'value': v, 'unit': 'MLUPS', 'cores': k, 'kind': 'port', 'single_core_value': v1,
               'sample': f'{args.cpu_size}x{args.cpu_size} periodic shear wave, {steps_cpu} steps after 1 warm-up, '
                         f'oracle/lbm_numpy.py (numpy restatement of the reference) on {k} processes x 1 thread with '
                         f'slab decomposition + ghost-row exchange (as mpirun -N {k}); host: {host_cores()} usable '
                         f'cores, {cpu_model()}; {dt:.1f} s; one process alone: {v1:.2f} MLUPS'}
    ref_cfg = None
    if rank == 0 and world == 1 and not args.no_ref_config and args.workload == 'shear':
        ref_cfg = reference_scaling_test()
    if rank == 0:
        line = {
            'metric': 'D2Q9 fp64 MLUPS', 'value': mlups, 'unit': 'MLUPS', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': dict(workload_config(args, world), **({'workload': f'von Karman rule set (inlet, outlet, plate) on {args.size}x{args.size}, bc_mode={args.bc_mode}', 'omega': float(np.reciprocal(3 * 0.04 + 0.5)), 'epsilon': None} if args.workload == 'karman' else {})), 'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
            'roofline': roofline, 'cpu_baseline': cpu, 'reference_scaling_test': ref_cfg,
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
    lat.load(hf, hr, hu, OMEGA)                       # H2D: 96 B per cell
    barrier()
    lat.run(K)
    for k in range(K):
        sink[...] = lat.probe_read(k + 1, 1)          # D2H: 16 B of every step, from the host-mapped ring as it arrives
    of, orho, ou = hf[g:NX - g], hr[g:NX - g], hu[g:NX - g]      # the rank's own rows (contiguous views)
    N.check(lib.lbm_materialize_region(lat._ctx, g, NX - g, 0, ny, N.dptr(of), N.dptr(orho), N.dptr(ou)))   # D2H: 96 B per cell
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    total_cells = nx_local * world * ny
    return {'value': total_cells * K / dt / 1e6, 'unit': 'MLUPS',
            'h2d_bytes_per_step': cells * 96.0 / K, 'd2h_bytes_per_step': cells * 96.0 / K + 16.0,
            'job': f'upload f,rho,u from pinned host memory ({cells * 96 / 1e9:.1f} GB per GPU), {K} steps enqueued at once, '
                   f'the 16-byte probe sample of every step read by the host from a host-mapped ring as the step completes, '
                   f'download f,rho,u; wall clock {dt:.2f} s, max over ranks',
            'steps': K}


if __name__ == '__main__':
    main()
