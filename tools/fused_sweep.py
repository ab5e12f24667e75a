"""Where does the two-steps-per-pass kernel pay? us per step of a periodic n x n lattice with one step per pass and
with two steps per pass at several segment lengths (rows per thread block). Run on a GPU box:

    python tools/fused_sweep.py [sizes...]
"""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from lattice_boltzmann_parallel_solver_b200.engine import Lattice


def time_steps(lat, n):
    st = torch.cuda.ExternalStream(lat.stream)
    lat.run(20)
    lat.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    lat.run(n)
    e1.record(st)
    lat.sync()
    return 1e3 * e0.elapsed_time(e1) / n


sizes = [int(a) for a in sys.argv[1:]] or [1024, 1536, 2048, 3072, 4096, 8192]
print(f'{"n":>6s} {"one step":>10s} ' + ' '.join(f'{"seg " + str(s):>10s}' for s in (8, 16, 32, 64, 128)) + '   (us per step; auto = library choice)')
for n in sizes:
    lat = Lattice(n, n)
    lat.load_equilibrium(1.0, ux_y=0.01 * np.sin(2 * np.pi * np.arange(n) / n))
    steps = max(40, min(400, int(2e9 / (n * n))))
    steps += steps % 2
    auto = time_steps(lat, steps)
    lat.set_option('fused', 0)
    row = [time_steps(lat, steps)]
    lat.set_option('fused', 1)
    for seg in (8, 16, 32, 64, 128):
        lat.set_option('fused_seg', seg)
        row.append(time_steps(lat, steps))
    lat.close()
    print(f'{n:6d} ' + ' '.join(f'{v:10.2f}' for v in row) + f'   auto {auto:.2f}  GLUPS best {n * n / min(row) / 1e3:.1f}')
