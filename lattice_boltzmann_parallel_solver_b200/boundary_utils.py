"""Drop-in for the reference's `src/boundary_utils.py`: the three scenario factories with the same signatures.

Each returns a `BoundaryBundle` — a callable with the reference's callback protocol
`boundary(f_pre_streaming, f_post_streaming, density, velocity, f_prev) -> f_post_streaming`
(src/lattice_boltzmann_method.py:222-223) that ALSO carries the ordered list of primitive overwrites, which
`lattice_boltzmann_step` folds into the fused kernel's per-cell kind map. Calling a bundle directly on arrays
applies the primitives one after another on the GPU, in the reference's order.
"""
import numpy as np

from . import boundary_spec as S
from .boundary_conditions import (BoundaryOp, inlet, moving_wall, outlet, periodic_with_pressure_variations,
                                  rigid_wall)
from .lattice_boltzmann_method import streaming
from .parallelization_utils import global_to_local_direction, x_in_process, y_in_process

INTERIOR = (slice(1, -1), slice(1, -1))


class BoundaryBundle:
    """Ordered (op, window) list. window None = whole array; INTERIOR = the [1:-1, 1:-1] view the parallel
    bundle applies inlet/outlet to (reference: src/boundary_utils.py:166-173)."""

    def __init__(self, name, shape=None):
        self.name = name
        self.shape = shape
        self.ops = []
        self.restream_after_first = False   # Poiseuille: stream again after the pressure BC (:103-104)
        self._compiled = {}

    def add(self, op, window=None):
        self.ops.append((op, window))
        return self

    def kind_map(self, shape):
        shape = (int(shape[0]), int(shape[1]))
        if self.shape is not None:
            assert shape == tuple(self.shape), f'{self.name}: built for {self.shape}, called with {shape}'
        if shape not in self._compiled:
            self._compiled[shape] = S.compile_ops(shape, self.ops)
        return self._compiled[shape]

    def __call__(self, f_pre_streaming, f_post_streaming, density=None, velocity=None, f_prev=None):
        for k, (op, window) in enumerate(self.ops):
            sx, sy = window if window is not None else (slice(None), slice(None))
            name = op.name
            if name == 'periodic_with_pressure_variations':
                f_pre_streaming = op(f_pre_streaming, density, velocity)
                if self.restream_after_first and k == 0:
                    f_post_streaming = streaming(f_pre_streaming)
            elif name == 'inlet':
                f_post_streaming[sx, sy, :] = op(np.ascontiguousarray(f_post_streaming[sx, sy, :]))
            elif name == 'outlet':
                f_post_streaming[sx, sy, :] = op(np.ascontiguousarray(f_prev[sx, sy, :]),
                                                 np.ascontiguousarray(f_post_streaming[sx, sy, :]))
            else:
                f_post_streaming = op(f_pre_streaming, f_post_streaming)
        return f_post_streaming


def _edge_mask(lx, ly, which):
    m = np.zeros((lx, ly), dtype=bool)
    if which == 'bottom':
        m[:, 0] = True
    elif which == 'top':
        m[:, -1] = True
    elif which == 'left':
        m[0, :] = True
    else:
        m[-1, :] = True
    return m


def couette_flow_boundary_conditions(lx: int, ly: int, U: float, avg_density: float) -> BoundaryBundle:
    """Rigid wall at y = ly-1, then a wall moving with (U, 0) at y = 0; periodic in x
    (reference: src/boundary_utils.py:9-55)."""
    b = BoundaryBundle('couette', (lx, ly))
    b.add(rigid_wall(_edge_mask(lx, ly, 'top')))
    b.add(moving_wall(_edge_mask(lx, ly, 'bottom'), np.array([U, 0]), avg_density))
    return b


def poiseuille_flow_boundary_conditions(lx: int, ly: int, p_in: float, p_out: float) -> BoundaryBundle:
    """Pressure-periodic virtual columns x = 0 / lx-1 (applied to f_pre, which is then streamed again), then
    rigid walls at y = 0 and y = ly-1 over all x (reference: src/boundary_utils.py:58-110)."""
    b = BoundaryBundle('poiseuille', (lx, ly))
    b.add(periodic_with_pressure_variations(_edge_mask(lx, ly, 'left') | _edge_mask(lx, ly, 'right'), p_in, p_out))
    b.restream_after_first = True
    b.add(rigid_wall(_edge_mask(lx, ly, 'bottom')))
    b.add(rigid_wall(_edge_mask(lx, ly, 'top')))
    return b


class _Bounce(BoundaryOp):
    """A bare list of bounce-back overwrites (the plate part of the parallel bundle writes them inline,
    reference: src/boundary_utils.py:178-201)."""
    name = 'plate'

    def __init__(self, shape):
        self.shape = shape
        self.items = []      # (index, dirs)

    def add(self, index, dirs):
        self.items.append((index, dirs))

    def emit(self, km):
        for index, dirs in self.items:
            km.bounce(index, dirs)

    def __call__(self, f_pre_streaming, f_post_streaming):
        return self._apply(self.shape, f_pre_streaming, f_post_streaming)


def parallel_von_karman_boundary_conditions(coord2d: list, n_local_x: int, n_local_y: int, lx: int, ly: int,
                                            x_size: int, y_size: int, density_in: float, velocity_in: float,
                                            plate_size: int) -> BoundaryBundle:
    """Inlet at global x = 0, outlet at x = lx-1 (copying from lx-2), thin plate between columns lx//4 and
    lx//4+1, all expressed on this rank's ghost-padded block (reference: src/boundary_utils.py:113-205).
    Raises NotImplementedError, like the reference (:174-176), when the last two columns sit on different ranks."""
    x_size, y_size = int(x_size), int(y_size)
    shape = (n_local_x + 2, n_local_y + 2)
    b = BoundaryBundle('von_karman', shape)
    if x_in_process(coord2d, 0, lx, x_size):
        b.add(inlet((n_local_x, n_local_y), density_in, velocity_in), INTERIOR)
    has_last, has_before_last = x_in_process(coord2d, lx - 1, lx, x_size), x_in_process(coord2d, lx - 2, lx, x_size)
    if has_last and has_before_last:
        b.add(outlet(), INTERIOR)
    elif has_last or has_before_last:
        raise NotImplementedError   # the reference's TODO: f_previous would have to be communicated

    y_lo_corner, y_hi_corner = ly // 2 - plate_size // 2, ly // 2 + plate_size // 2 - 1
    full = [global_to_local_direction(coord2d[1], y, ly, y_size) for y in range(y_lo_corner + 1, y_hi_corner)
            if y_in_process(coord2d, y, ly, y_size)]
    plate = _Bounce(shape)
    for gx, dirs, top_dirs, bottom_dirs in ((lx // 4, [1, 5, 8], [1, 8], [1, 5]),
                                            (lx // 4 + 1, [3, 6, 7], [3, 7], [3, 6])):
        if not x_in_process(coord2d, gx, lx, x_size):
            continue
        x = global_to_local_direction(coord2d[0], gx, lx, x_size)
        if full:
            plate.add((x, np.array(full)), dirs)
        if y_in_process(coord2d, y_hi_corner, ly, y_size):
            plate.add((x, global_to_local_direction(coord2d[1], y_hi_corner, ly, y_size)), top_dirs)
        if y_in_process(coord2d, y_lo_corner, ly, y_size):
            plate.add((x, global_to_local_direction(coord2d[1], y_lo_corner, ly, y_size)), bottom_dirs)
    if plate.items:
        b.add(plate)
    return b
