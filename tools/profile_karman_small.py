"""Workload for an ncu capture of ONE step kernel of config 4 (von Karman 422 x 182 with ghost ring, mask kernel with ghost
stores, `k_step<1,1,0,0>`), launched one by one (LBM_NO_GRAPHS=1 in the environment):

    LBM_NO_GRAPHS=1 ncu --set full --clock-control none -k regex:k_step -s 20 -c 1 -o out python tools/profile_karman_small.py
"""
import sys
sys.path.insert(0, '.')
import lattice_boltzmann_parallel_solver_b200 as P
from lattice_boltzmann_parallel_solver_b200.engine import Lattice
from oracle import lbm_numpy as onp

BU = P.boundary_utils
rho, u = onp.uniform((422, 182), 1.0, 0.1, 0.0)
bc = BU.parallel_von_karman_boundary_conditions([0, 0], 420, 180, 420, 180, 1, 1, 1.0, 0.1, 40)
lat = Lattice(422, 182, bc.kind_map((422, 182)), ghost=(1, 1))
lat.connect_self_periodic()
lat.load(onp.equilibrium(rho, u), rho, u, 1.6)
lat.run(40)
lat.sync()
lat.close()
