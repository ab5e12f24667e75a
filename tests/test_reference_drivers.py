"""BASELINE.json's north star: "experiments.py and main.py drive it unchanged". The reference's OWN driver files
(copied unmodified to baseline/_ref/src by `__graft_entry__.build()`; matplotlib / pygifsicle are recording stubs under
tests/stubs, mpi4py the package's single-process stand-in) are executed twice by tests/ref_drivers.py — once entirely on
the reference's numpy modules, once with this package's drop-in modules first on the path — and everything they
produce must be IDENTICAL: the viscosity-vs-omega curves (whole-field amin / amax after every step, omega sweep over
one initial state, src/experiments.py:147-223), the Couette profile and regression csv (:304-374), the probe trace of
x_strouhal (:650-720), scaling_test (:723-774), the kept-and-read-later lines of plot_evolution_of_velocity (:102-144)
and `main.py -f couette_vectors` end to end (src/main.py:36-223).

CPU: the drop-in modules run on tests/fake_native.py (host logic: handles, deferral, boundary compilation, drivers).
GPU (`-m gpu`): the same drivers on the CUDA library."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, 'tests', 'ref_drivers.py')


def _have_reference():
    return any(os.path.exists(os.path.join(d, 'experiments.py'))
               for d in (os.path.join(ROOT, 'baseline', '_ref', 'src'), '/root/reference/src'))


def _run(mode, out):
    res = subprocess.run([sys.executable, WORKER, mode, out], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0 and f'OK {mode}' in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
    return np.load(out)


def _compare(ref, got, what):
    assert sorted(got.files) == sorted(ref.files)
    for k in got.files:
        assert ref[k].shape == got[k].shape, (what, k, ref[k].shape, got[k].shape)
        assert np.array_equal(ref[k], got[k], equal_nan=True), \
            f'{what}: {k} differs from the reference run (max abs diff {np.nanmax(np.abs(ref[k] - got[k])):.3e})'
    assert ref['vel_at_p'].shape == (91,) and ref['evolution_lines'].shape == (11, 16) and ref['couette_csv'].shape == (5,)
    assert ref['couette_evolution'].shape == (10, 12) and ref['poiseuille_evolution'].shape == (10, 10)


@pytest.fixture(scope='module')
def reference_outputs(tmp_path_factory):
    if not _have_reference():
        pytest.skip('reference drivers not installed (baseline/_ref/src is filled by __graft_entry__.build())')
    return _run('reference', str(tmp_path_factory.mktemp('ref') / 'reference.npz'))


def test_reference_drivers_run_unchanged_on_the_dropin_modules_cpu(reference_outputs, tmp_path):
    _compare(reference_outputs, _run('fake', str(tmp_path / 'fake.npz')), 'fake native')


@pytest.mark.gpu
def test_reference_drivers_run_unchanged_on_the_gpu(reference_outputs, tmp_path):
    _compare(reference_outputs, _run('gpu', str(tmp_path / 'gpu.npz')), 'CUDA library')
