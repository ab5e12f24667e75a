"""Workload for an ncu capture of the cluster kernel: config 1 (100 x 50 periodic shear wave), 2000 steps in one launch."""
import sys
sys.path.insert(0, '.')
from lattice_boltzmann_parallel_solver_b200.engine import Lattice
from oracle import lbm_numpy as onp

rho, u = onp.sinusoidal_velocity_x((100, 50), 0.01)
lat = Lattice(100, 50)
lat.set_option('cluster', 2)
lat.load(onp.equilibrium(rho, u), rho, u, 1.0)
lat.run(2000)
lat.sync()
lat.close()
