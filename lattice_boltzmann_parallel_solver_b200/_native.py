"""ctypes binding of include/lbm_b200.h (lattice_boltzmann_parallel_solver_b200/liblbm_b200.so).

There is no fallback: if the CUDA library is missing or no GPU is visible, every compute entry point raises
`LbmNativeError`. Error codes are mapped to the exception types the reference raises for the same
conditions: LBM_ERR_ARG -> AssertionError (the reference's asserts,
src/lattice_boltzmann_method.py:210-212, src/boundary_conditions.py:89,105-106), everything else ->
RuntimeError subclasses.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LBM_B200_LIB') or os.path.join(HERE, 'liblbm_b200.so')

LBM_OK, LBM_ERR_ARG, LBM_ERR_CUDA, LBM_ERR_STATE, LBM_ERR_NOMEM, LBM_ERR_TIMEOUT = range(6)

RULE_PULL, RULE_BOUNCE, RULE_CONST, RULE_OUTLET = range(4)
CELL_OUTLET_SRC, CELL_PBC_IN_SRC, CELL_PBC_OUT_SRC = 1, 2, 4
BC_AUTO, BC_MASK, BC_EDGE = range(3)


class LbmNativeError(RuntimeError):
    pass


class LbmStateError(LbmNativeError):
    pass


class LbmTimeoutError(LbmNativeError):
    pass


class Kind(C.Structure):
    _fields_ = [('rule', C.c_uint8 * 9), ('flags', C.c_uint8), ('skip_store', C.c_uint16)]


KIND_DTYPE = np.dtype([('rule', np.uint8, (9,)), ('flags', np.uint8), ('skip_store', np.uint16)])
assert KIND_DTYPE.itemsize == C.sizeof(Kind) == 12


class BcDesc(C.Structure):
    _fields_ = [('n_kinds', C.c_int), ('kinds', C.POINTER(Kind)),
                ('n_k_rows', C.c_int), ('k_table', C.POINTER(C.c_double)),
                ('n_c_rows', C.c_int), ('c_table', C.POINTER(C.c_double)),
                ('pbc_rho_in', C.c_double), ('pbc_rho_out', C.c_double),
                ('kind_map', C.POINTER(C.c_uint8))]


class HaloExport(C.Structure):
    _fields_ = [('mem_handle', C.c_uint8 * 64), ('device', C.c_int32), ('nx', C.c_int32), ('ny', C.c_int32),
                ('pitch', C.c_int32), ('pid', C.c_int64), ('arena_ptr', C.c_uint64), ('arena_bytes', C.c_int64)]


_DP = C.POINTER(C.c_double)
_CTX = C.c_void_p

# name -> (restype, argtypes); every symbol include/lbm_b200.h declares
SIGNATURES = {
    'lbm_last_error': (C.c_char_p, []),
    'lbm_version': (C.c_char_p, []),
    'lbm_device_count': (C.c_int, []),
    'lbm_equilibrium': (C.c_int, [C.c_int, C.c_int64, _DP, _DP, _DP]),
    'lbm_density': (C.c_int, [C.c_int, C.c_int64, _DP, _DP]),
    'lbm_velocity': (C.c_int, [C.c_int, C.c_int64, _DP, _DP, _DP]),
    'lbm_streaming': (C.c_int, [C.c_int, C.c_int, C.c_int, _DP, _DP]),
    'lbm_selftest_arith': (C.c_int, [C.c_int, C.c_int64, C.c_uint64, C.POINTER(C.c_uint64)]),
    'lbm_plan_two_step': (C.c_int, [C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]),
    'lbm_bc_apply': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(BcDesc), _DP, _DP, _DP]),
    'lbm_pbc_apply': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _DP, _DP, _DP]),
    'lbm_create': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(BcDesc), C.POINTER(_CTX)]),
    'lbm_destroy': (C.c_int, [_CTX]),
    'lbm_set_bc_mode': (C.c_int, [_CTX, C.c_int]),
    'lbm_set_option': (C.c_int, [_CTX, C.c_char_p, C.c_int]),
    'lbm_device_bytes': (C.c_int64, [_CTX]),
    'lbm_stream': (C.c_void_p, [_CTX]),
    'lbm_upload': (C.c_int, [_CTX, _DP, _DP, _DP, C.c_double]),
    'lbm_init_equilibrium': (C.c_int, [_CTX, _DP, _DP, C.c_double, C.c_double, C.c_double, C.c_double]),
    'lbm_step': (C.c_int, [_CTX, C.c_double, C.c_int]),
    'lbm_run_host': (C.c_int, [_CTX, _DP, _DP, _DP, C.c_double, C.c_int, _DP, _DP, _DP]),
    'lbm_sync': (C.c_int, [_CTX]),
    'lbm_time': (C.c_int64, [_CTX]),
    'lbm_launch_count': (C.c_int64, [_CTX]),
    'lbm_materialize': (C.c_int, [_CTX, _DP, _DP, _DP]),
    'lbm_materialize_region': (C.c_int, [_CTX, C.c_int, C.c_int, C.c_int, C.c_int, _DP, _DP, _DP]),
    'lbm_probe_config': (C.c_int, [_CTX, C.c_int, C.c_int, C.c_int]),
    'lbm_probe_read': (C.c_int, [_CTX, C.c_int64, C.c_int, _DP]),
    'lbm_minmax': (C.c_int, [_CTX, C.c_int, C.c_int, C.c_int, C.c_int, _DP]),
    'lbm_history_config': (C.c_int, [_CTX, C.c_int]),
    'lbm_history_store': (C.c_int, [_CTX, C.c_int]),
    'lbm_history_read': (C.c_int, [_CTX, C.c_int, _DP, _DP]),
    'lbm_halo_export_handle': (C.c_int, [_CTX, C.POINTER(HaloExport)]),
    'lbm_halo_connect': (C.c_int, [_CTX, C.c_int, C.POINTER(HaloExport)]),
    'lbm_halo_finalize': (C.c_int, [_CTX]),
}

_lib = None


def load():
    """Loads the shared library (no CUDA call is made). Raises LbmNativeError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LbmNativeError(
                f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a). There is no CPU fallback.')
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc == LBM_OK:
        return
    msg = load().lbm_last_error().decode('utf-8', 'replace')
    if rc == LBM_ERR_ARG:
        raise AssertionError(msg)
    if rc == LBM_ERR_STATE:
        raise LbmStateError(msg)
    if rc == LBM_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == LBM_ERR_TIMEOUT:
        raise LbmTimeoutError(msg)
    raise LbmNativeError(msg)


_device = None


def device():
    """The CUDA device this process computes on: LOCAL_RANK under torchrun, else 0. Raises without a GPU."""
    global _device
    if _device is None:
        n = load().lbm_device_count()
        if n <= 0:
            raise LbmNativeError('no CUDA device is visible: this package has no CPU fallback')
        _device = int(os.environ.get('LBM_DEVICE', os.environ.get('LOCAL_RANK', '0'))) % n
    return _device


def set_device(index):
    global _device
    _device = int(index)


def dptr(a):
    return a.ctypes.data_as(_DP) if a is not None else None


def as_f64(a, shape=None, what='array'):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), f'{what}: expected shape {tuple(shape)}, got {a.shape}'
    return a
