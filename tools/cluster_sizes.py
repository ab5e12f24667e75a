"""Cluster kernel on configs 1-3 with 8 and 16 CTAs per cluster (LBM_CLUSTER_SIZE is read when a context plans its
cluster, so every size runs in its own process): us per step of 20 000 steps in one launch + parity against the C oracle.

    python tools/cluster_sizes.py            # both sizes
    python tools/cluster_sizes.py worker     # the size in the environment
"""
import os
import subprocess
import sys
sys.path.insert(0, '.')


def worker():
    import numpy as np
    import torch
    import lattice_boltzmann_parallel_solver_b200 as P
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    from oracle import lbm_c as oc, lbm_numpy as onp
    BU = P.boundary_utils
    rng = np.random.default_rng(0)
    cases = (('config1 periodic 100x50', (100, 50), None, oc.periodic(), 1.0),
             ('config2 couette 100x100', (100, 100), lambda s: BU.couette_flow_boundary_conditions(*s, 0.05, 1.0).kind_map(s), oc.couette(0.05, 1.0), 1.0),
             ('config3 poiseuille 100x50', (100, 50), lambda s: BU.poiseuille_flow_boundary_conditions(*s, 0.3338, 0.3328).kind_map(s), oc.poiseuille(0.3338, 0.3328), 1.5),
             ('periodic 37x23', (37, 23), None, oc.periodic(), 1.2),
             ('periodic 160x120', (160, 120), None, oc.periodic(), 1.2))
    for name, shp, mk, scen, om in cases:
        r = rng.uniform(0.9, 1.1, shp); u = rng.uniform(-0.05, 0.05, shp + (2,)); f = onp.equilibrium(r, u)
        lat = Lattice(*shp, mk(shp) if mk else None)
        lat.set_option('cluster', 2)
        lat.load(f, r, u, om)
        k0 = lat.launches
        lat.run(33)
        ok = lat.launches - k0 == 1 and all(np.array_equal(a, b) for a, b in zip(lat.fields(), oc.run(f, r, u, om, scen, 33)))
        st = torch.cuda.ExternalStream(lat.stream)
        lat.run(2000); lat.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0 = lat.launches
        e0.record(st); lat.run(20000); e1.record(st); lat.sync()
        print(f"cluster of {os.environ.get('LBM_CLUSTER_SIZE', 'default'):>7s}  {name:28s} {1e3 * e0.elapsed_time(e1) / 20000:6.2f} us/step in "
              f"{lat.launches - k0} launch(es), 33 steps equal the oracle: {ok}", flush=True)
        lat.close()


if __name__ == '__main__':
    if sys.argv[1:] == ['worker']:
        worker()
    else:
        for size, mlim in (('8', None), ('16', None), (None, '512')):
            env = dict(os.environ)
            env.pop('LBM_CLUSTER_SIZE', None)
            env.pop('LBM_CLUSTER_MLIM', None)
            if size:
                env['LBM_CLUSTER_SIZE'] = size
            if mlim:
                env['LBM_CLUSTER_MLIM'] = mlim     # cells per CTA above which a thread takes two cells
                print(f'-- two cells per thread above {mlim} cells per CTA', flush=True)
            subprocess.run([sys.executable, __file__, 'worker'], env=env, check=False)
