"""Runs the REFERENCE's own test files (tests/*.py of the reference, copied unmodified to baseline/_ref/tests by
`__graft_entry__.build()`) inside a scratch tree laid out like the reference repository — `src/` + `tests/`, run from its
root with `src` and `tests` on PYTHONPATH, as the reference's Makefile does (Makefile:3, :13-14):

    python tests/ref_own_tests.py tree     <dir>    build the tree: src/ = the reference's modules with the four hot-path
                                                    modules replaced by this package's drop-in shims; tests/ = its tests
    python tests/ref_own_tests.py fixtures <dir>    f_0.npy .. f_10.npy of tests/test_parallelization_von_karman.py:56-66 (the
                                                    reference repository does not hold them): its loop (:24-55) on the
                                                    REFERENCE's own modules, one rank
    python tests/ref_own_tests.py unit   fake|gpu|reference <dir>    `pytest tests` (the 18 unittest cases) in the tree
    python tests/ref_own_tests.py karman fake|gpu <dir>              tests/test_parallelization_von_karman.py as a script
                                                                     (under torchrun: k ranks, the reference's k-rank grid)
Own process for every mode: the flat module names must not leak into the test session.
"""
import os
import runpy
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'lattice_boltzmann_parallel_solver_b200')
REF = os.path.join(ROOT, 'baseline', '_ref')
HOT = ('lattice_boltzmann_method.py', 'boundary_conditions.py', 'boundary_utils.py', 'parallelization_utils.py')


def make_tree(tree):
    """<tree>/reference: the reference as it is; <tree>/dropin: the four hot-path modules replaced by the drop-in shims."""
    for flavour in ('reference', 'dropin'):
        os.makedirs(os.path.join(tree, flavour, 'src'), exist_ok=True)
        for name in os.listdir(os.path.join(REF, 'src')):
            shim = flavour == 'dropin' and name in HOT
            shutil.copyfile(os.path.join(PKG, 'dropin', name) if shim else os.path.join(REF, 'src', name),
                            os.path.join(tree, flavour, 'src', name))
        shutil.copytree(os.path.join(REF, 'tests'), os.path.join(tree, flavour, 'tests'), dirs_exist_ok=True)


def set_path(tree, mode):
    """-> the repository-like root the mode runs in (cwd; `src` and `tests` on the path as in the reference's Makefile)."""
    root = os.path.join(tree, 'reference' if mode == 'reference' else 'dropin')
    sys.path[:0] = [os.path.join(root, 'src'), os.path.join(root, 'tests'), root, os.path.join(ROOT, 'tests', 'stubs'),
                    os.path.join(PKG, 'dropin'), ROOT]
    if mode == 'fake':
        from lattice_boltzmann_parallel_solver_b200 import _native as N
        from tests.fake_native import FakeLib
        fake = FakeLib()
        N.load = lambda: fake
        N.device = lambda: 0
    os.chdir(root)
    return root


def fixtures(tree):
    root = set_path(tree, 'reference')
    import numpy as np
    from mpi4py import MPI
    from boundary_utils import parallel_von_karman_boundary_conditions
    from initial_values import density_1_velocity_x_u0_velocity_y_0_initial
    from lattice_boltzmann_method import equilibrium_distr_func, lattice_boltzmann_step
    from parallelization_utils import communication
    import lattice_boltzmann_method
    assert os.path.dirname(os.path.abspath(lattice_boltzmann_method.__file__)) == os.path.join(root, 'src')
    lx, ly, d, u0 = 420, 180, 40, 0.1
    omega = np.reciprocal(3 * 0.04 + 0.5)
    cart = MPI.COMM_WORLD.Create_cart(dims=[1, 1], periods=[True, True], reorder=False)
    density, velocity = density_1_velocity_x_u0_velocity_y_0_initial((lx + 2, ly + 2), u0)
    f = equilibrium_distr_func(density, velocity)
    bound = parallel_von_karman_boundary_conditions([0, 0], lx, ly, lx, ly, 1, 1, 1.0, u0, d)
    com = communication(cart)
    for i in range(11):
        f, density, velocity = lattice_boltzmann_step(f, density, velocity, omega, bound, com)
        for flavour in ('reference', 'dropin'):
            np.save(os.path.join(tree, flavour, 'tests', 'von_karman_vortex_shedding', f'f_{i}.npy'), f[1:-1, 1:-1, :])
    print('OK fixtures', flush=True)


def unit(mode, tree):
    root = set_path(tree, mode)
    import pytest
    args = ['tests', '-q', '-p', 'no:cacheprovider', '-s']
    rc = pytest.main(args)
    import src.lattice_boltzmann_method as tested       # what the reference's test files import
    assert os.path.dirname(os.path.abspath(tested.__file__)) == os.path.join(root, 'src'), tested.__file__
    if mode != 'reference':
        assert tested.lattice_boltzmann_step.__module__.startswith('lattice_boltzmann_parallel_solver_b200'), tested.lattice_boltzmann_step.__module__
    print(f'{"OK" if rc == 0 else "FAILED"} unit {mode}', flush=True)
    sys.exit(int(rc))


def karman(mode, tree):
    root = set_path(tree, mode)
    runpy.run_path(os.path.join(root, 'tests', 'test_parallelization_von_karman.py'), run_name='__main__')
    from lattice_boltzmann_parallel_solver_b200 import lattice_boltzmann_method as impl
    impl.release_lattices()
    print(f'OK karman {mode} rank {os.environ.get("RANK", "0")} of {os.environ.get("WORLD_SIZE", "1")}', flush=True)


if __name__ == '__main__':
    what = sys.argv[1]
    if what == 'tree':
        make_tree(sys.argv[2])
    elif what == 'fixtures':
        fixtures(sys.argv[2])
    elif what == 'unit':
        unit(sys.argv[2], sys.argv[3])
    elif what == 'karman':
        karman(sys.argv[2], sys.argv[3])
