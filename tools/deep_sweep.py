"""A/B of the multi-step kernels on the headline lattice: us per time step of an n x n periodic shear-wave lattice for
k_step2x (depth 2, 24-slot ring), k_stepNx<2|3|4> (18-slot rings) at several segment lengths and L2-prefetch distances.
Passes are pure (fused_exact): n_steps is a multiple of the depth. Run on a GPU box:

    python tools/deep_sweep.py [n] [--quick]
"""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from lattice_boltzmann_parallel_solver_b200.engine import Lattice


def time_steps(lat, n, reps=2):
    st = torch.cuda.ExternalStream(lat.stream)
    lat.run(12)
    lat.sync()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        lat.run(n)
        e1.record(st)
        lat.sync()
        best = min(best, 1e3 * e0.elapsed_time(e1) / n)
    return best


args = [a for a in sys.argv[1:] if not a.startswith('--')]
n = int(args[0]) if args else 16384
quick = '--quick' in sys.argv
PFS = [int(v) for a in sys.argv for v in (a[5:].split(',') if a.startswith('--pf=') else [])]
lat = Lattice(n, n)
lat.load_equilibrium(1.0, ux_y=0.01 * np.sin(2 * np.pi * np.arange(n) / n))
lat.set_option('fused_exact', 1)
steps = 48
lat.set_option('fused', 0)
t1 = time_steps(lat, 24)
print(f'{n}x{n}: one step per pass {t1:9.1f} us/step  {n * n / t1 / 1e3:7.1f} GLUPS', flush=True)
lat.set_option('fused', 1)
configs = [('k_step2x', 2, 0), ('k_stepNx<2>', 2, 1), ('k_stepNx<3>', 3, 1), ('k_stepNx<4>', 4, 1)]
if '--d3' in sys.argv:
    configs = [('k_stepNx<3>', 3, 1)]
if '--d23' in sys.argv:
    configs = [('k_stepNx<2>', 2, 1), ('k_stepNx<3>', 3, 1)]
for name, depth, deep2 in configs:
    lat.set_option('fused_depth', depth)
    lat.set_option('deep2', deep2)
    for seg in ((128,) if quick else (32, 64, 128, 256, 0)):       # 0 = the library's choice (wave-aware)
        lat.set_option('fused_seg', seg)
        for pf in (PFS if PFS else ((2,) if quick else (0, 2, 4))):
            lat.set_option('l2_prefetch', pf)
            t = time_steps(lat, steps)
            print(f'{name:12s} seg {seg:4d} pf {pf}: {t:9.1f} us/step  {n * n / t / 1e3:7.1f} GLUPS', flush=True)
lat.close()
