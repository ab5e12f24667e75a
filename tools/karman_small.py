"""Config 4 alone (von Karman 420 x 180 + ghost ring, the reference's scaling_test lattice) for profiling:

    LBM_NO_GRAPHS=1 ncu --set full --import-source on -k regex:k_step -s 60 -c 1 -o out python tools/karman_small.py 100
"""
import sys
sys.path.insert(0, '.')
import numpy as np
import lattice_boltzmann_parallel_solver_b200 as P
from lattice_boltzmann_parallel_solver_b200.engine import Lattice

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rho = np.ones((422, 182))
u = np.zeros((422, 182, 2))
u[..., 0] = 0.1
f = P.lattice_boltzmann_method.equilibrium_distr_func(rho, u)
bc = P.boundary_utils.parallel_von_karman_boundary_conditions([0, 0], 420, 180, 420, 180, 1, 1, 1.0, 0.1, 40)
lat = Lattice(422, 182, bc.kind_map((422, 182)), ghost=(1, 1))
lat.connect_self_periodic()
lat.load(f, rho, u, 1.6)
lat.run(steps)
lat.sync()
print('ran', steps, 'steps,', lat.launches, 'launches')
lat.close()
