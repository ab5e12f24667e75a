import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import lattice_boltzmann_parallel_solver_b200 as P
from lattice_boltzmann_parallel_solver_b200.engine import Lattice
from oracle import lbm_numpy as onp
BU = P.boundary_utils
def timeit(name, lat, n):
    st = torch.cuda.ExternalStream(lat.stream)
    lat.run(200); lat.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(st); lat.run(n); t1 = time.perf_counter(); e1.record(st); lat.sync(); t2 = time.perf_counter()
    print(f'{name:34s} {n} steps: enqueue {1e6*(t1-t0)/n:6.2f} us/step, wall {1e6*(t2-t0)/n:6.2f} us/step, gpu {1e3*e0.elapsed_time(e1)/n:6.2f} us/step, {lat.nx*lat.ny*n/(t2-t0)/1e6:9.1f} MLUPS')
rho,u = onp.sinusoidal_velocity_x((100,50),0.01); f = onp.equilibrium(rho,u)
lat = Lattice(100,50); lat.load(f,rho,u,1.0); timeit('config1 shear 100x50', lat, 20000); lat.close()
rho,u = onp.uniform((100,100)); f = onp.equilibrium(rho,u)
lat = Lattice(100,100, BU.couette_flow_boundary_conditions(100,100,0.05,1.0).kind_map((100,100))); lat.load(f,rho,u,1.0); timeit('config2 couette 100x100', lat, 20000); lat.close()
rho,u = onp.uniform((100,50)); f = onp.equilibrium(rho,u)
lat = Lattice(100,50, BU.poiseuille_flow_boundary_conditions(100,50,0.3338,0.3328).kind_map((100,50))); lat.load(f,rho,u,1.5); timeit('config3 poiseuille 100x50', lat, 20000); lat.close()
rho,u = onp.uniform((422,182),1.0,0.1,0.0); f = onp.equilibrium(rho,u)
bc = BU.parallel_von_karman_boundary_conditions([0,0],420,180,420,180,1,1,1.0,0.1,40)
lat = Lattice(422,182,bc.kind_map((422,182)),ghost=(1,1)); lat.connect_self_periodic(); lat.load(f,rho,u,1.6); timeit('config4 karman 422x182 (ghost)', lat, 20000)
lat.probe(316, 91, 65536); timeit('config4 karman + probe', lat, 20000); lat.close()
rho,u = onp.uniform((420,180),1.0,0.1,0.0); f = onp.equilibrium(rho,u)
lat = Lattice(420,180); lat.load(f,rho,u,1.6); timeit('periodic 420x180 (generic)', lat, 20000); lat.close()
lat = Lattice(1024,1024); lat.load_equilibrium(1.0); timeit('periodic 1024x1024 (pair)', lat, 5000); lat.close()
