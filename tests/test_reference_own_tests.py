"""The reference's OWN test suite (its tests/*.py, copied unmodified to baseline/_ref/tests by `__graft_entry__.build()`)
run against this package's drop-in modules, in a scratch tree laid out like the reference repository and with the
reference's own invocation (`pytest tests` from the root, `src` and `tests` on PYTHONPATH; Makefile:3, :13-14):

* the 18 unittest cases of test_{boundary_conditions,density_computation,initial_values,mass_preservation,
  navier_stokes_eq,streaming_func,velocity_computation}.py (`from src.lattice_boltzmann_method import ...`);
* tests/test_parallelization_von_karman.py as the script it is (`mpirun -n k python tests/...` in the reference; here
  `torchrun` with k ranks): the probe trace against the reference's own fixture vel_at_p.npy with `==`, and the gathered
  populations of 11 steps (through `save_mpiio`) against f_i.npy (:56-66). The reference repository does not hold the f_i
  fixtures; tests/ref_own_tests.py generates them by running that loop on the reference's own modules.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, 'tests', 'ref_own_tests.py')


def _run(*args, timeout=900, env=None):
    res = subprocess.run([sys.executable, WORKER] + list(args), capture_output=True, text=True, timeout=timeout, cwd=ROOT,
                         env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


@pytest.fixture(scope='module')
def tree(tmp_path_factory):
    if not os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'tests', 'test_mass_preservation.py')):
        pytest.skip("the reference's tests are not installed (baseline/_ref/tests is filled by __graft_entry__.build())")
    t = str(tmp_path_factory.mktemp('reference_repo'))
    _run('tree', t)
    out = _run('unit', 'reference', t)                 # the suite as the reference runs it: the expectation
    assert '18 passed' in out and 'OK unit reference' in out, out[-2000:]
    assert 'OK fixtures' in _run('fixtures', t)
    return t


def test_reference_unit_tests_on_the_dropin_modules_cpu(tree):
    out = _run('unit', 'fake', tree)
    assert '18 passed' in out and 'OK unit fake' in out, out[-2000:]


def test_reference_parallel_von_karman_test_on_the_dropin_modules_cpu(tree):
    assert 'OK karman fake rank 0 of 1' in _run('karman', 'fake', tree)


@pytest.mark.gpu
def test_reference_unit_tests_on_the_gpu(tree):
    out = _run('unit', 'gpu', tree)
    assert '18 passed' in out and 'OK unit gpu' in out, out[-2000:]


def _gpu_count():
    from lattice_boltzmann_parallel_solver_b200 import _native
    return int(_native.load().lbm_device_count())


@pytest.mark.gpu
@pytest.mark.parametrize('size', [1, 2, 4, 6])
def test_reference_parallel_von_karman_test_on_the_gpu(tree, size):
    """k ranks on the reference's k-rank grid (1x1, 1x2, 2x2, 2x3), one process per GPU where the box has k GPUs, otherwise
    all ranks on GPU 0 (gloo plumbing; the halo itself is CUDA-IPC peer stores either way)."""
    if size == 1:
        assert 'OK karman gpu rank 0 of 1' in _run('karman', 'gpu', tree)
        return
    env = dict(os.environ, LBM_HALO_TIMEOUT_S='120')
    if _gpu_count() < size:
        env['LBM_DIST_BACKEND'] = 'gloo'
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={size}', '--master-addr',
           '127.0.0.1', '--master-port', str(29570 + size), WORKER, 'karman', 'gpu', tree]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    for r in range(size):
        assert f'OK karman gpu rank {r} of {size}' in res.stdout
