"""Workload for an ncu capture of the cluster kernel, 2000 steps in one launch:

    python tools/profile_cluster.py            config 1 (100 x 50 periodic shear wave)
    python tools/profile_cluster.py couette    config 2 (100 x 100, moving + rigid wall)
    python tools/profile_cluster.py poiseuille config 3 (100 x 50, pressure-periodic + walls)
"""
import sys
sys.path.insert(0, '.')
import lattice_boltzmann_parallel_solver_b200 as P
from lattice_boltzmann_parallel_solver_b200.engine import Lattice
from oracle import lbm_numpy as onp

BU = P.boundary_utils
case = sys.argv[1] if len(sys.argv) > 1 else 'periodic'
if case == 'couette':
    shape, omega = (100, 100), 1.0
    rho, u = onp.uniform(shape)
    lat = Lattice(*shape, BU.couette_flow_boundary_conditions(*shape, 0.05, 1.0).kind_map(shape))
elif case == 'poiseuille':
    shape, omega = (100, 50), 1.5
    rho, u = onp.uniform(shape)
    lat = Lattice(*shape, BU.poiseuille_flow_boundary_conditions(*shape, 0.3338, 0.3328).kind_map(shape))
else:
    shape, omega = (100, 50), 1.0
    rho, u = onp.sinusoidal_velocity_x(shape, 0.01)
    lat = Lattice(*shape)
lat.set_option('cluster', 2)
lat.load(onp.equilibrium(rho, u), rho, u, omega)
lat.run(2000)
lat.sync()
lat.close()
