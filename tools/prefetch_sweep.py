"""A/B of the two-step kernel's L2 prefetch distance (rows ahead) on one lattice, back to back in one process:

    python tools/prefetch_sweep.py [n] [steps]
"""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from lattice_boltzmann_parallel_solver_b200.engine import Lattice

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
lat = Lattice(n, n)
lat.load_equilibrium(1.0, ux_y=0.01 * np.sin(2 * np.pi * np.arange(n) / n))
st = torch.cuda.ExternalStream(lat.stream)
lat.run(40)
lat.sync()
for pf in (0, 2, 1, 3, 4, 6, 0, 2):
    lat.set_option('l2_prefetch', pf)
    lat.run(10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    lat.run(steps)
    e1.record(st)
    lat.sync()
    ms = e0.elapsed_time(e1) / steps
    print(f'n {n} l2_prefetch {pf}: {ms:.4f} ms/step  {n * n / ms / 1e3:.0f} MLUPS', flush=True)
lat.close()
