"""Workload for ncu captures of the multi-step kernels: python tools/profile_deep.py <depth> <seg> [n] [passes]"""
import sys
sys.path.insert(0, '.')
import numpy as np
from lattice_boltzmann_parallel_solver_b200.engine import Lattice

depth, seg = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
passes = int(sys.argv[4]) if len(sys.argv) > 4 else 3
lat = Lattice(n, n)
lat.load_equilibrium(1.0, ux_y=0.01 * np.sin(2 * np.pi * np.arange(n) / n))
lat.set_option('fused_exact', 1)
lat.set_option('fused_depth', depth)
lat.set_option('deep2', 1)
lat.set_option('fused_seg', seg)
lat.run(depth * passes)
lat.sync()
lat.close()
