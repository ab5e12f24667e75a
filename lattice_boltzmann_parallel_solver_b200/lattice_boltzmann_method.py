"""Drop-in for the reference's `src/lattice_boltzmann_method.py` — same names, arguments, return values and
assertion behaviour; every array operation runs on the GPU through the C-ABI of include/lbm_b200.h.

`lattice_boltzmann_step` keeps the reference's functional contract (inputs are never mutated, three results
per call) but its results are lazy, device-resident `LatticeArray` handles (engine.py); passing them back in
— the loop every driver of the reference runs — advances the resident lattice by one fused kernel launch.
"""
import atexit
from typing import Callable, Tuple

import numpy as np

from . import _native as N


def get_velocity_sets() -> np.ndarray:
    """D2Q9 velocity set c_i, shape (9, 2) (reference: src/lattice_boltzmann_method.py:5-26)."""
    cx = (0, 1, 0, -1, 0, 1, -1, -1, 1)
    cy = (0, 0, 1, 0, -1, 1, 1, -1, -1)
    return np.stack([np.array(cx), np.array(cy)], axis=1)


def vel_to_opp_vel_mapping() -> np.ndarray:
    """Index of the opposite direction (reference: src/lattice_boltzmann_method.py:29-39)."""
    return np.array([0, 3, 4, 1, 2, 7, 8, 5, 6])


def get_w_i() -> np.ndarray:
    """D2Q9 weights (reference: src/lattice_boltzmann_method.py:42-52)."""
    return np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4)


def reynolds_number(L: int, u: float, v: float) -> float:
    """Re = L u / v (reference: src/lattice_boltzmann_method.py:55-71) — scalar host arithmetic."""
    return np.divide(L * u, v)


def strouhal_number(f: float, L: int, u: float) -> float:
    """St = f L / u (reference: src/lattice_boltzmann_method.py:74-90) — scalar host arithmetic."""
    return np.divide(f * L, u)


def compute_density(prob_densitiy_func: np.ndarray) -> np.ndarray:
    """rho = sum_i f_i in numpy's pairwise order (reference: src/lattice_boltzmann_method.py:93-105)."""
    assert prob_densitiy_func.shape[-1] == 9
    f = N.as_f64(prob_densitiy_func)
    out = np.empty(f.shape[:-1])
    N.check(N.load().lbm_density(N.device(), out.size, N.dptr(f), N.dptr(out)))
    return out


def compute_velocity_field(density_func: np.ndarray, prob_density_func: np.ndarray) -> np.ndarray:
    """u = sum_i c_i f_i / rho, zero where rho == 0 (reference: src/lattice_boltzmann_method.py:108-137)."""
    assert prob_density_func.shape[-1] == 9
    f = N.as_f64(prob_density_func)
    rho = N.as_f64(density_func, f.shape[:-1], 'density')
    out = np.empty(rho.shape + (2,))
    N.check(N.load().lbm_velocity(N.device(), rho.size, N.dptr(rho), N.dptr(f), N.dptr(out)))
    return out


def streaming(prob_density_func: np.ndarray) -> np.ndarray:
    """Periodic streaming, out[x, y, i] = in[(x, y) - c_i, i] (reference: src/lattice_boltzmann_method.py:140-159)."""
    assert prob_density_func.shape[-1] == 9
    f = N.as_f64(prob_density_func)
    assert f.ndim == 3
    out = np.empty_like(f)
    N.check(N.load().lbm_streaming(N.device(), f.shape[0], f.shape[1], N.dptr(f), N.dptr(out)))
    return out


def equilibrium_distr_func(density_func: np.ndarray, velocity_field: np.ndarray) -> np.ndarray:
    """f_eq_i = w_i rho (1 + 3 c.u + 4.5 (c.u)^2 - 1.5 |u|^2) in the reference's rounding order
    (reference: src/lattice_boltzmann_method.py:162-188). Accepts any leading shape, like the reference."""
    density_func = np.asarray(density_func)
    velocity_field = np.asarray(velocity_field)
    assert density_func.shape == velocity_field.shape[:-1]
    rho = N.as_f64(density_func)
    u = N.as_f64(velocity_field)
    lead = rho.shape if rho.ndim >= 2 else (1,) + rho.shape   # 1-D inputs give (1, n, 9) there (matmul broadcasting)
    out = np.empty(lead + (9,))
    N.check(N.load().lbm_equilibrium(N.device(), rho.size, N.dptr(rho), N.dptr(u), N.dptr(out)))
    return out


# ---------------------------------------------------------------------------------------------------------
# the time step
# ---------------------------------------------------------------------------------------------------------
_lattices = {}   # (nx, ny, id(boundary), id(comm)) -> Lattice; lattices are reused across uploads


def _resolve_boundary(boundary, shape):
    if boundary is None:
        return None
    compiler = getattr(boundary, 'kind_map', None)
    if compiler is None:
        raise TypeError(
            'boundary must come from this package\'s boundary_utils factories (or be a BoundaryBundle): an arbitrary '
            'Python closure cannot be folded into the CUDA kernel and there is no CPU fallback')
    return compiler(shape)


def _lattice_for(shape, boundary, comm):
    from .engine import Lattice
    key = (shape[0], shape[1], id(boundary), id(comm))
    hit = _lattices.get(key)
    if hit is not None and hit[1] is boundary and hit[2] is comm:
        return hit[0]
    km = _resolve_boundary(boundary, shape)
    ghost = (1, 1) if comm is not None else (0, 0)
    lat = Lattice(shape[0], shape[1], km, ghost)
    if comm is not None:
        attach = getattr(comm, 'attach', None)
        if attach is None:
            raise TypeError('parallel_communication must come from this package\'s parallelization_utils.communication')
        attach(lat)
    if len(_lattices) > 8:     # a sweep over sizes should not pin device memory forever
        _lattices.pop(next(iter(_lattices)))[0].retire()   # results still referenced are brought to the host first
    _lattices[key] = (lat, boundary, comm)
    return lat


def release_lattices():
    """Frees every cached device lattice (results still referenced are brought to the host first)."""
    while _lattices:
        _lattices.popitem()[1][0].retire()


def flush_all():
    """Launch every queued step of every lattice that has halo neighbours. Called before this package's blocking
    process-group operations (dist.WorldComm.Barrier / allgather / Sendrecv) and at interpreter exit: a neighbouring
    rank's kernel of the same step waits for this rank's, so a rank must not block on the host with steps queued."""
    for lat, _, comm in list(_lattices.values()):
        if comm is not None and lat._pending_n:
            lat.flush()


def _flush_at_exit():
    # A rank whose loop ended with a deferred step must still launch it: neighbouring ranks' kernels of the same
    # step wait for its "begun" flag (include/lbm_b200.h, halo section).
    for lat, _, comm in list(_lattices.values()):
        if comm is not None and lat._pending_n:
            try:
                lat.flush()
                lat.sync()
            except Exception:
                pass


atexit.register(_flush_at_exit)


def lattice_boltzmann_step(f: np.ndarray, density: np.ndarray, velocity: np.ndarray, omega: float,
                           boundary: Callable = None,
                           parallel_communication: Callable = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One BGK time step: collide with the GIVEN moments, halo exchange, stream, boundary, new moments
    (reference: src/lattice_boltzmann_method.py:191-228). Returns (f, density, velocity) of the next time as
    lazy device-resident arrays; the arguments are not modified."""
    assert tuple(f.shape[0:2]) == tuple(density.shape)
    assert tuple(f.shape[0:2]) == tuple(velocity.shape[0:2])
    assert 0 < omega < 2
    shape = (int(f.shape[0]), int(f.shape[1]))
    lat = _lattice_for(shape, boundary, parallel_communication)
    if not lat.is_current(f, density, velocity):
        lat.reset_for_upload()
        if parallel_communication is not None:
            parallel_communication.before_load(lat)
        lat.load(np.asarray(f), np.asarray(density), np.asarray(velocity), omega)
        if parallel_communication is not None:
            parallel_communication.after_load(lat)
    return lat.request_step(omega)
