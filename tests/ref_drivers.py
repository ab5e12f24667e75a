"""Runs the REFERENCE's own, unmodified drivers (src/experiments.py, src/main.py; copied to baseline/_ref/src by
`__graft_entry__.build()`) and dumps what they produced, in one of three configurations:

    python tests/ref_drivers.py reference <out.npz>   everything from the reference (its numpy time step): the expectation
    python tests/ref_drivers.py fake      <out.npz>   this package's drop-in modules on tests/fake_native.py (CPU)
    python tests/ref_drivers.py gpu       <out.npz>   this package's drop-in modules on the CUDA library

matplotlib / pygifsicle are test stubs (tests/stubs); mpi4py is the package's single-process stand-in. The drivers run
in a scratch directory with the `figures/` tree they write into. Own process: the flat module names (`experiments`,
`lattice_boltzmann_method`, ...) must not leak into the test session.
"""
import contextlib
import io
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'lattice_boltzmann_parallel_solver_b200')


def ref_src():
    for cand in (os.path.join(ROOT, 'baseline', '_ref', 'src'), '/root/reference/src'):
        if os.path.exists(os.path.join(cand, 'experiments.py')):
            return cand
    return None


def main():
    mode, out = sys.argv[1], sys.argv[2]
    src = ref_src()
    assert src, 'reference drivers not available (baseline/_ref/src is filled by __graft_entry__.build())'
    stubs, dropin = os.path.join(ROOT, 'tests', 'stubs'), os.path.join(PKG, 'dropin')
    # reference: its own modules first (dropin only supplies `mpi4py`); ours: the drop-in modules shadow the reference's
    sys.path[:0] = [src, stubs, dropin, ROOT] if mode == 'reference' else [dropin, stubs, src, ROOT]
    if mode == 'fake':
        from lattice_boltzmann_parallel_solver_b200 import _native as N
        from tests.fake_native import FakeLib
        fake = FakeLib()
        N.load = lambda: fake
        N.device = lambda: 0
    work = tempfile.mkdtemp(prefix='lbm_ref_drivers_')
    for d in ('shear_wave_decay', 'couette_flow', 'poiseuille_flow', 'von_karman_vortex_shedding/reynold_strouhal',
              'von_karman_vortex_shedding/nx_strouhal', 'von_karman_vortex_shedding/blockage_strouhal',
              'von_karman_vortex_shedding/scaling_test', 'von_karman_vortex_shedding/all_png_parallel'):
        os.makedirs(os.path.join(work, 'figures', d))
    os.chdir(work)
    import matplotlib
    import experiments as E
    import lattice_boltzmann_method as L
    where = os.path.dirname(os.path.abspath(L.__file__))
    assert where == (src if mode == 'reference' else dropin), f'wrong lattice_boltzmann_method on the path: {where}'
    res = {}
    quiet = contextlib.redirect_stderr(io.StringIO())   # tqdm bars

    def plotted(prefix):
        return [a for name, a, k in matplotlib.calls if name.startswith(prefix)]

    with quiet:
        # (1) shear-wave viscosity vs omega: whole-field np.amin / np.amax after EVERY step, omega sweep over one initial
        #     state (src/experiments.py:147-223)
        matplotlib.calls.clear()
        E.plot_measured_viscosity_vs_omega(lattice_grid_shape=(24, 20), time_steps=260, omega_discretization=3)
        curves = [a for a in plotted('plt.subplots()[') if len(a) == 2]
        for i, (x, y) in enumerate(curves):
            res[f'visc_{i}_x'], res[f'visc_{i}_y'] = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
        assert len(curves) == 4, len(curves)
        # (2) Couette: moving wall + rigid wall, profile + linear regression written to csv (:304-374)
        matplotlib.calls.clear()
        E.plot_couette_flow_vel_vectors(lattice_grid_shape=(12, 14), omega=1.0, U=0.05, time_steps=150)
        res['couette_csv'] = np.genfromtxt('figures/couette_flow/linregress.csv', delimiter=',', skip_header=1)
        res['couette_profile'] = np.asarray(plotted('plt.plot')[0][0], dtype=float)
        # (3) x_strouhal: parallel von Karman path on one rank, one velocity cell read after every step (:650-720)
        E.x_strouhal('reynold_strouhal', lattice_grid_shape=(60, 36), plate_size=10, time_steps=90)
        res['vel_at_p'] = np.load('figures/von_karman_vortex_shedding/reynold_strouhal/vel_at_p_25.npy')
        # (4) scaling_test: the reference's published benchmark loop (:723-774)
        E.scaling_test('scaling_test', lattice_grid_shape=(60, 36), plate_size=10, time_steps=70)
        res['scaling_mlups_finite'] = np.array([float(np.isfinite(np.load(
            'figures/von_karman_vortex_shedding/scaling_test/60_36_1.npy')[0]))])
        # (5) velocity evolution: keeps results of many steps and looks at them after the loop (:102-144)
        matplotlib.calls.clear()
        E.plot_evolution_of_velocity(lattice_grid_shape=(20, 16), epsilon=0.05, omega=1.2, time_steps=100, number_of_visualizations=10)
        lines = [a for a in plotted('plt.subplots()[') if len(a) >= 2 and np.ndim(a[1]) == 1 and len(a[1]) == 16]
        res['evolution_lines'] = np.array([np.asarray(a[1], dtype=float) for a in lines])
        # (7), (8) Couette / Poiseuille evolution: velocities.append(velocity) after EVERY step, a few of them plotted after
        #     the loop (:225-301, :503-583) -- kept results are parked in the device history, not copied out per step
        for tag, call, ly in (('couette_evolution', lambda: E.plot_couette_flow_evolution(
                                  lattice_grid_shape=(14, 12), omega=1.1, U=0.04, time_steps=90, number_of_visualizations=10), 12),
                              ('poiseuille_evolution', lambda: E.plot_poiseuille_flow_evolution(
                                  lattice_grid_shape=(18, 10), omega=1.4, delta_p=0.002, time_steps=90, number_of_visualizations=10), 10)):
            matplotlib.calls.clear()
            call()
            sim = [a for name, a, k in matplotlib.calls if name.endswith('.plot') and k.get('linestyle') == ':']
            assert len(sim) == 10, len(sim)
            res[tag] = np.array([np.asarray(a[0], dtype=float) for a in sim])
            assert res[tag].shape == (10, ly)
        # (9) density evolution: a density column plotted every few steps (:55-99)
        matplotlib.calls.clear()
        E.plot_evolution_of_density(lattice_grid_shape=(18, 18), initial_p0=0.5, epsilon=0.08, omega=1.0, time_steps=100,
                                    number_of_visualizations=10)
        cols = [a for name, a, k in matplotlib.calls if name.endswith('.plot') and len(a) == 2 and np.ndim(a[1]) == 1 and len(a[1]) == 18]
        res['density_evolution'] = np.array([np.asarray(a[1], dtype=float) for a in cols])
        assert res['density_evolution'].shape[0] >= 10
        # (10) Poiseuille profiles, area under the curve, pressure along the centre line, absolute error (:377-504)
        matplotlib.calls.clear()
        E.plot_poiseuille_flow_vel_vectors(lattice_grid_shape=(24, 12), omega=1.5, delta_p=0.002, time_steps=300)
        lines = [np.asarray(a[0], dtype=float) for name, a, k in matplotlib.calls
                 if name == 'plt.plot' and len(a) >= 2 and np.ndim(a[0]) == 1 and np.asarray(a[0]).dtype.kind == 'f']
        res['poiseuille_vectors'] = np.concatenate([v.ravel() for v in lines])
        assert len(lines) >= 3
        # (11) the parallel von Karman driver on one rank: a velocity-magnitude frame every 100 steps through save_mpiio (:584-647)
        matplotlib.calls.clear()
        if mode == 'reference':     # the reference's save_mpiio needs MPI-IO (Cartcomm.Sub, MPI.File): on ONE rank it is np.save
            E.save_mpiio = lambda comm, file_name, array: np.save(file_name, array)
        E.plot_parallel_von_karman_vortex_street(lattice_grid_shape=(60, 36), plate_size=10, time_steps=201)
        frames = [a[0] for name, a, k in matplotlib.calls if name == 'cm.viridis']
        res['karman_frames'] = np.array(frames)
        assert res['karman_frames'].shape == (3, 36, 60)
    # (6) main.py end to end (argparse -> experiments): couette_vectors with explicit sizes
    matplotlib.calls.clear()
    sys.argv = ['main.py', '-f', 'couette_vectors', '-l', '10', '12', '-t', '60', '-mwv', '0.03']
    import main as M
    with quiet:
        M.main()
    res['main_couette_csv'] = np.genfromtxt('figures/couette_flow/linregress.csv', delimiter=',', skip_header=1)
    if mode != 'reference':
        from lattice_boltzmann_parallel_solver_b200 import lattice_boltzmann_method as impl
        impl.release_lattices()
    np.savez(out, **res)
    print('OK', mode, sorted(res), flush=True)


if __name__ == '__main__':
    main()
