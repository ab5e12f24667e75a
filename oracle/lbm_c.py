"""ctypes binding of the C oracle (oracle/lbm_oracle.c) — TEST INFRASTRUCTURE, NOT PRODUCT.

Same import restriction as `oracle/lbm_numpy.py`. Builds `oracle/liblbm_oracle.so` with gcc on first use if
it is missing (the GPU box receives the prebuilt file). Scenario parameters are derived here exactly as the
reference's factories derive them (citations inline).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import lbm_numpy as onp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'liblbm_oracle.so')
_lib = None


class Scenario(C.Structure):
    _fields_ = [('kind', C.c_int), ('k_mov', C.c_double * 3), ('rho_in', C.c_double), ('rho_out', C.c_double),
                ('cin', C.c_double * 9), ('plate_x', C.c_int), ('y_full0', C.c_int), ('y_full1', C.c_int),
                ('y_top', C.c_int), ('y_bot', C.c_int), ('probe_x', C.c_int), ('probe_y', C.c_int)]


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, 'lbm_oracle.c')):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'liblbm_oracle.so'], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        _lib.orc_run.argtypes = [C.c_int, C.c_int, dp, dp, dp, C.c_double, C.POINTER(Scenario), C.c_int, dp]
        _lib.orc_run.restype = C.c_int
        _lib.orc_equilibrium.argtypes = [C.c_long, dp, dp, dp]
        _lib.orc_density.argtypes = [C.c_long, dp, dp]
        _lib.orc_velocity.argtypes = [C.c_long, dp, dp, dp]
        _lib.orc_stream.argtypes = [C.c_int, C.c_int, dp, dp]
        _lib.orc_collide.argtypes = [C.c_long, dp, dp, dp, C.c_double, dp]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def periodic():
    s = Scenario()
    s.kind = 0
    s.probe_x = s.probe_y = -1
    return s


def couette(U, avg_density):
    """src/boundary_utils.py:34-38 + src/boundary_conditions.py:207-210."""
    s = periodic()
    s.kind = 1
    k = onp.moving_wall_constants((4, 7, 8), np.array([U, 0]), avg_density)
    s.k_mov[:] = [k[4], k[7], k[8]]
    return s


def poiseuille(p_in, p_out):
    """src/boundary_conditions.py:304-309."""
    s = periodic()
    s.kind = 2
    s.rho_in = float(np.divide(p_in, onp.CS2))
    s.rho_out = float(np.divide(p_out, onp.CS2))
    return s


def karman(lx, ly, rho_in, u_in, plate, ghost, probe=None):
    """Plate geometry of src/boundary_utils.py:145-149,184-201 (ghost=1: parallel path on one rank) or
    milestoneQuickFunctionCalls.py:309-311 (ghost=0: serial rigid_object)."""
    s = periodic()
    s.kind = 3 if ghost else 4
    s.cin[:] = list(onp.inlet_constants(rho_in, u_in))
    s.plate_x = lx // 4 + ghost
    s.y_full0 = ly // 2 - plate // 2 + 1 + ghost
    s.y_full1 = ly // 2 + plate // 2 - 1 + ghost
    s.y_top = ly // 2 + plate // 2 - 1 + ghost
    s.y_bot = ly // 2 - plate // 2 + ghost
    if probe is not None:
        s.probe_x, s.probe_y = probe[0] + ghost, probe[1] + ghost
    return s


def run(f, rho, u, omega, scenario, n_steps, want_probe=False):
    """Advance copies of (f, rho, u) by n_steps; returns (f, rho, u[, probe (n_steps, 2)])."""
    f = np.array(f, dtype=np.float64, order='C')
    rho = np.array(rho, dtype=np.float64, order='C')
    u = np.array(u, dtype=np.float64, order='C')
    nx, ny = rho.shape
    probe = np.zeros((n_steps, 2)) if want_probe else None
    rc = lib().orc_run(nx, ny, _p(f), _p(rho), _p(u), float(omega), C.byref(scenario), int(n_steps),
                       _p(probe) if want_probe else None)
    assert rc == 0
    return (f, rho, u, probe) if want_probe else (f, rho, u)


def equilibrium(rho, u):
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(rho.shape + (9,))
    lib().orc_equilibrium(rho.size, _p(rho), _p(u), _p(out))
    return out
