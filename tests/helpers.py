"""Shared helpers for the parity tests."""
import hashlib

import numpy as np

# per field max|a-b|/max|b| <= 1e-12 AND allclose(rtol=1e-12, atol=1e-12*max|b|)
RTOL = 1e-12


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rel_err(a, b):
    m = float(np.max(np.abs(b)))
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) / (m if m > 0 else 1.0)


def assert_parity(a, b, what='', exact=True):
    """Bit-exact is the target (all operations are IEEE add/mul/div/sqrt in a fixed order); the documented
    bar of BASELINE.json's north_star is 1e-12 relative. `exact=True` asserts the former."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f'{what}: shape {a.shape} vs {b.shape}'
    if exact:
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            extent = ', '.join(f'axis{k}: {int(bad[:, k].min())}..{int(bad[:, k].max())} ({len(np.unique(bad[:, k]))} distinct)'
                               for k in range(bad.shape[1]))
            raise AssertionError(f'{what}: {len(bad)} of {a.size} values differ bitwise; first at {bad[0]}: '
                                 f'{a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}; rel err {rel_err(a, b):.3e}; where: {extent}')
    else:
        m = float(np.max(np.abs(b)))
        assert rel_err(a, b) <= RTOL, f'{what}: rel err {rel_err(a, b):.3e}'
        assert np.allclose(a, b, rtol=RTOL, atol=RTOL * m), what
