"""Device-resident lattice + the lazy result handles `lattice_boltzmann_step` returns.

`Lattice` is a thin object over one `lbm_ctx` of the C-ABI (include/lbm_b200.h). `LatticeArray` is what the
drop-in `lattice_boltzmann_step` hands back in place of the reference's three fresh numpy arrays
(src/lattice_boltzmann_method.py:225-228): an ndarray-like handle on "f / density / velocity at time t of this
lattice" which is only copied to the host when somebody looks at it. Feeding the three handles straight back
into the next call — what every driver loop of the reference does — costs one kernel launch and
no transfer.

Safety rule: the device keeps S_{t-1} and S_t (A/B buffers); values of time t are reconstructed from S_{t-1}.
Before the device advances past t, every handle of time t that is still referenced anywhere is materialised
(weak references tell). To make that cheap in the common loop `f, rho, u = step(f, rho, u, ...)`, steps are
DEFERRED: a call only queues its step (up to `MAX_DEFERRED`, same omega) and hands out the handles of the future
time; the queue is launched when somebody looks at a result, when it is full, or when omega changes. By then the
loop has dropped the old handles, and a whole batch goes to `lbm_step(n)` at once — which is what lets the
reference's own driver loops run on the two-steps-per-pass kernel and on CUDA-graph replay.

Drivers that look at ONE velocity cell after every step (`vel_at_p.append(np.linalg.norm(velocity[px, py, ...]))`,
experiments.py:703-704) cannot defer anything; for them the second consecutive read of the same cell configures the
device-side probe on it (`Lattice._probe_sample`), after which a read costs one step launch and a 16-byte copy out
of the host-mapped probe ring.

Drivers that KEEP a field of every step and look at a few afterwards (`velocities.append(velocity)`,
experiments.py:254, :542): when the device has to move past a time whose density / velocity handles are still
referenced, those fields are parked in a device history slot by one asynchronous launch (`Lattice._park`,
lbm_history_*) and only come to the host if somebody reads them; with no free slot (or when `f` itself is kept) they
are materialised to the host as before.
"""
import os
import weakref

import numpy as np

from . import _native as N

HISTORY_BYTES = int(os.environ.get('LBM_HISTORY_MB', '256')) << 20   # device memory for kept-but-unread (rho, u) fields
HISTORY_MAX_SLOTS = 1024
# steps queued by lattice_boltzmann_step before they are launched as one batch: a multiple of the three steps of a pass
# (64 left one slow single-step launch per batch on bandwidth-bound lattices) and of the 32 steps of a replayed graph
MAX_DEFERRED = 96
AUTO_PROBE_CAPACITY = 4096   # ring entries of the probe that velocity[px, py] reads configure by themselves


class Lattice:
    """One device-resident lattice (one `lbm_ctx`). Native API for callers that do not need the reference's
    per-step Python protocol: `load`, `run(n)`, `fields()`, `probe`, `minmax`."""

    def __init__(self, nx, ny, kind_map=None, ghost=(0, 0), device=None, bc_mode=N.BC_AUTO):
        self.lib = N.load()
        self.device = N.device() if device is None else int(device)
        self.nx, self.ny = int(nx), int(ny)
        self.ghost = (int(ghost[0]), int(ghost[1]))
        self._ctx = N._CTX()
        desc_ref = None
        if kind_map is not None and not kind_map.is_trivial:
            assert kind_map.shape == (self.nx, self.ny)
            desc, self._keep = kind_map.to_desc()
            desc_ref = N.C.byref(desc)
        N.check(self.lib.lbm_create(self.device, self.nx, self.ny, self.ghost[0], self.ghost[1], desc_ref,
                                    N.C.byref(self._ctx)))
        self._keep = None
        if bc_mode != N.BC_AUTO:
            N.check(self.lib.lbm_set_bc_mode(self._ctx, bc_mode))
        self.omega = None
        self.time = 0                 # reference steps taken on the device since load
        self._probe = None            # (x, y, capacity) of the configured probe
        self._probe_t0 = 0            # samples exist for device times > _probe_t0
        self._probe_auto = False      # configured by _probe_sample, not by the caller
        self._watch = None            # (x, y, t) of the last velocity cell read
        self._probe_buf = None        # (array, pointer) of _probe_cell
        self._shapes = {k: fn(self.nx, self.ny) for k, fn in _SHAPES.items()}
        # lazy-handle bookkeeping (see module docstring)
        self._pending = None          # omega of the steps requested but not yet launched
        self._pending_n = 0           # how many of them
        self._generation = 0          # bumped by every load: handles of an earlier upload never count as current
        self._handles = {}            # api time -> list of weakrefs to LatticeArray
        # device-side history (lbm_history_*): results that are kept but not looked at are parked on the device
        self._hist_free = None        # free slot indices; None = not configured yet, [] may also mean "switched off"
        self._parked = weakref.WeakSet()

    # ---- lifetime ----------------------------------------------------------------------------------------
    def retire(self):
        """Free the device lattice but keep every result that is still referenced readable (materialise first)."""
        if self._ctx:
            try:
                self.reset_for_upload()
            finally:
                self.close()

    def close(self):
        if self._ctx:
            for slot in list(self._parked):   # results parked on the device must outlive it
                slot.drain()
            self.lib.lbm_destroy(self._ctx)
            self._ctx = N._CTX()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- native API --------------------------------------------------------------------------------------
    @property
    def shape(self):
        return (self.nx, self.ny)

    @property
    def stream(self):
        """cudaStream_t (int) the step kernels run on — wrap with torch.cuda.ExternalStream for event timing."""
        return int(self.lib.lbm_stream(self._ctx) or 0)

    @property
    def device_bytes(self):
        return int(self.lib.lbm_device_bytes(self._ctx))

    @property
    def launches(self):
        return int(self.lib.lbm_launch_count(self._ctx))

    def set_option(self, name, value):
        """'fused' (two steps per pass), 'graphs' (CUDA-graph replay), 'generic_kernel' — see include/lbm_b200.h."""
        N.check(self.lib.lbm_set_option(self._ctx, name.encode(), int(value)))

    def load(self, f, rho, u, omega):
        """State triple of the reference (f, density, velocity) -> device; first collision with the given moments."""
        f = N.as_f64(f, (self.nx, self.ny, 9), 'f')
        rho = N.as_f64(rho, (self.nx, self.ny), 'density')
        u = N.as_f64(u, (self.nx, self.ny, 2), 'velocity')
        assert 0 < omega < 2
        N.check(self.lib.lbm_upload(self._ctx, N.dptr(f), N.dptr(rho), N.dptr(u), float(omega)))
        self.omega, self.time = float(omega), 0
        self._probe_t0, self._watch = 0, None
        self._generation += 1

    def load_equilibrium(self, omega, rho_x=None, ux_y=None, rho0=1.0, ux0=0.0, uy0=0.0):
        """f = f_eq(rho, u) of a separable initial field, built on the device (initial_values.py:38-123)."""
        rx = N.as_f64(rho_x, (self.nx,), 'rho_x') if rho_x is not None else None
        uy = N.as_f64(ux_y, (self.ny,), 'ux_y') if ux_y is not None else None
        assert 0 < omega < 2
        N.check(self.lib.lbm_init_equilibrium(self._ctx, N.dptr(rx), N.dptr(uy), float(rho0), float(ux0), float(uy0),
                                              float(omega)))
        self.omega, self.time = float(omega), 0
        self._probe_t0, self._watch = 0, None
        self._generation += 1

    def run(self, n_steps, omega=None):
        """n reference time steps, asynchronous."""
        omega = self.omega if omega is None else float(omega)
        assert 0 < omega < 2
        N.check(self.lib.lbm_step(self._ctx, omega, int(n_steps)))
        self.omega = omega
        self.time += int(n_steps)

    def run_host(self, f, rho, u, omega, n_steps, out=None):
        """The whole job from and to host arrays: load(f, rho, u, omega) + run(n_steps) + fields(), same results — on a
        fluid lattice without ghosts upload, time steps and download are pipelined over row chunks (lbm_run_host).
        `out` = (f_out, density_out, velocity_out) arrays to fill (any may be None; they may be the inputs themselves);
        default: fresh arrays. Returns the three output arrays."""
        f = N.as_f64(f, (self.nx, self.ny, 9), 'f')
        rho = N.as_f64(rho, (self.nx, self.ny), 'density')
        u = N.as_f64(u, (self.nx, self.ny, 2), 'velocity')
        assert 0 < omega < 2 and n_steps >= 1
        self.reset_for_upload()           # queued steps are launched, results still referenced are brought to the host
        if out is None:
            out = (np.empty_like(f), np.empty_like(rho), np.empty_like(u))
        for a, shape in zip(out, ((self.nx, self.ny, 9), (self.nx, self.ny), (self.nx, self.ny, 2))):
            assert a is None or (a.shape == shape and a.dtype == np.float64 and a.flags.c_contiguous)
        N.check(self.lib.lbm_run_host(self._ctx, N.dptr(f), N.dptr(rho), N.dptr(u), float(omega), int(n_steps),
                                      N.dptr(out[0]), N.dptr(out[1]), N.dptr(out[2])))
        self.omega, self.time = float(omega), int(n_steps)
        self._probe_t0, self._watch = 0, None
        self._generation += 1
        return out

    def sync(self):
        N.check(self.lib.lbm_sync(self._ctx))

    def fields(self, f=True, rho=True, u=True, region=None):
        """Reference-layout (f_post, density, velocity) of the current time; None for the ones not asked for."""
        x0, x1, y0, y1 = (0, self.nx, 0, self.ny) if region is None else region
        n = (x1 - x0, y1 - y0)
        of = np.empty(n + (9,)) if f else None
        orho = np.empty(n) if rho else None
        ou = np.empty(n + (2,)) if u else None
        N.check(self.lib.lbm_materialize_region(self._ctx, x0, x1, y0, y1, N.dptr(of), N.dptr(orho), N.dptr(ou)))
        return of, orho, ou

    def probe(self, x, y, capacity=1 << 16):
        """Record (u_x, u_y) at one cell after every step (experiments.py:703-704)."""
        N.check(self.lib.lbm_probe_config(self._ctx, int(x), int(y), int(capacity)))
        self._probe = (int(x), int(y), int(capacity))
        self._probe_t0, self._probe_auto = self.time, False

    def probe_read(self, t0, n):
        out = np.empty((n, 2))
        N.check(self.lib.lbm_probe_read(self._ctx, int(t0), int(n), N.dptr(out)))
        return out

    def _probe_cell(self, t):
        """One sample through a preallocated buffer (the per-step read of a driver loop: no allocation, no pointer cast)."""
        buf = self._probe_buf
        if buf is None:
            arr = np.empty((1, 2))
            buf = self._probe_buf = (arr, N.dptr(arr))
        rc = self.lib.lbm_probe_read(self._ctx, t, 1, buf[1])
        if rc:
            N.check(rc)
        return buf[0][0].copy()

    def _probe_sample(self, x, y, t):
        """(u_x, u_y) of cell (x, y) at the current device time t from the probe ring, or None when the ring does not
        hold it. The reference's drivers read ONE velocity cell after every step (experiments.py:703-704): the second
        read of the same cell at consecutive times configures the probe on it, and from then on such a read is a
        16-byte copy out of host-mapped memory instead of a materialisation launch. A probe the caller configured
        is used when it sits on the cell, and never replaced."""
        gx, gy = self.ghost
        if not (gx <= x < self.nx - gx and gy <= y < self.ny - gy) or t != self.time:
            return None                               # ghost cells are not computed by the step kernels
        p = self._probe
        if p is not None and p[0] == x and p[1] == y and t > self._probe_t0:
            return self._probe_cell(t)
        if (p is None or self._probe_auto) and self._watch == (x, y, t - 1):
            self.probe(x, y, capacity=AUTO_PROBE_CAPACITY)
            self._probe_auto = True
        self._watch = (x, y, t)
        return None

    def minmax(self, region=None):
        """(min rho, max rho, min u, max u) of the current state, reduced on the device (experiments.py:181-193)."""
        x0, x1, y0, y1 = (0, self.nx, 0, self.ny) if region is None else region
        out = np.empty(4)
        N.check(self.lib.lbm_minmax(self._ctx, x0, x1, y0, y1, N.dptr(out)))
        return tuple(out)

    def halo_export(self):
        e = N.HaloExport()
        N.check(self.lib.lbm_halo_export_handle(self._ctx, N.C.byref(e)))
        return e

    def halo_connect(self, slot, export):
        N.check(self.lib.lbm_halo_connect(self._ctx, int(slot), N.C.byref(export)))

    def halo_finalize(self):
        N.check(self.lib.lbm_halo_finalize(self._ctx))

    def connect_self_periodic(self):
        """One rank: every neighbour is this lattice itself (communication() with a 1x1 topology)."""
        e = self.halo_export()
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                if (dx or dy) and (not dx or self.ghost[0]) and (not dy or self.ghost[1]):
                    self.halo_connect((dx + 1) * 3 + dy + 1, e)
        self.halo_finalize()

    # ---- lazy-handle protocol used by lattice_boltzmann_step ------------------------------------------------
    @property
    def api_time(self):
        """Time of the newest handles handed out (device time + queued steps)."""
        return self.time + self._pending_n

    def _live(self, t):
        refs = self._handles.get(t, ())
        return [h for h in (r() for r in refs) if h is not None]

    def _preserve(self, t):
        """Materialise every still-referenced handle of device time t (must be the current device time)."""
        live = [h for h in self._live(t) if h._value is None and h._hist is None]
        if live and t == self.time and t > 0:
            want = {h._which for h in live}
            if 'f' not in want and self._park(live):
                self._handles.pop(t, None)
                return
            f, rho, u = self.fields('f' in want, 'rho' in want, 'u' in want)
            for h in live:
                h._value = {'f': f, 'rho': rho, 'u': u}[h._which]
                h._value.setflags(write=False)
        self._handles.pop(t, None)

    def _park(self, handles):
        """Keep the density / velocity of the current time in a device history slot for these handles (one asynchronous
        launch) instead of bringing them to the host now: the `velocities.append(velocity)` loops of the reference
        (experiments.py:254, :542) look at a handful of the fields they keep. False when no slot is to be had."""
        if self._hist_free is None:
            self._hist_free = []
            per_slot = self.nx * self.ny * 24
            n = min(HISTORY_MAX_SLOTS, HISTORY_BYTES // per_slot)
            if n >= 4 and self.ghost[0] < 2:
                try:
                    N.check(self.lib.lbm_history_config(self._ctx, int(n)))
                    self._hist_free = list(range(int(n) - 1, -1, -1))
                except MemoryError:
                    pass
        if not self._hist_free:
            return False
        slot = _HistorySlot(self, self._hist_free.pop(), handles)
        N.check(self.lib.lbm_history_store(self._ctx, slot.index))
        return True

    def flush(self, upto=None):
        """Launch the queued steps (all of them, or up to api time `upto`). Times whose handles are still
        referenced are stopped at on the way, so that those handles can be materialised before the device moves on."""
        goal = self.api_time if upto is None else min(int(upto), self.api_time)
        while self.time < goal:
            self._preserve(self.time)
            stop = goal
            if goal - self.time > 1:                  # (a single step has no time in between to stop at)
                for t in sorted(self._handles):
                    if self.time < t < goal and any(h._value is None for h in self._live(t)):
                        stop = t
                        break
            n = stop - self.time
            omega = self._pending
            self._pending_n -= n
            if self._pending_n == 0:
                self._pending = None
            self.run(n, omega)
        if len(self._handles) > 1:
            for t in [t for t in self._handles if t < self.time and not self._live(t)]:
                del self._handles[t]

    def request_step(self, omega):
        """Called by lattice_boltzmann_step: queues one step and returns the three handles of its result."""
        omega = float(omega)
        if self._pending_n and (omega != self._pending or self._pending_n >= MAX_DEFERRED):
            self.flush()
        self._pending = omega
        self._pending_n += 1
        t = self.api_time
        ref = weakref.ref
        hs = (LatticeArray(self, t, 'f'), LatticeArray(self, t, 'rho'), LatticeArray(self, t, 'u'))
        self._handles[t] = [ref(hs[0]), ref(hs[1]), ref(hs[2])]
        return hs

    def reset_for_upload(self):
        """Before a fresh upload: nothing pending may be lost, nothing referenced may go stale."""
        self.flush()
        self._preserve(self.time)
        self._handles.clear()

    def is_current(self, *handles):
        t, g = self.time + self._pending_n, self._generation
        for h in handles:
            if not (type(h) is LatticeArray and h._lattice is self and h._generation == g and h._t == t and not h._dirty):
                return False
        return True


def connect_blocks(blocks, dims):
    """Wire a whole periodic Cartesian topology of lattices that live in THIS process (one per block coordinate,
    possibly on different GPUs): blocks[(cx, cy)] -> Lattice, dims = (x_size, y_size). The multi-process equivalent
    is `parallelization_utils.communication(comm).attach(lattice)`.

    Stepping such blocks: a block's step kernel WAITS (spinning on a flag) until its neighbours have finished the
    previous step. One process per GPU, that is always safe. Several blocks on ONE device are only safe when a waiting
    kernel can never keep the kernel it waits for from running: step the blocks with `run_blocks` (which drains every
    launch group before the next when blocks share a device), or launch them yourself one step at a time with a
    `sync()` of all blocks in between. Queueing many steps per block asynchronously on a shared device can dead-lock;
    the wait then times out (LBM_HALO_TIMEOUT_S) and every later call raises LbmTimeoutError — it never returns wrong
    fields."""
    xs, ys = int(dims[0]), int(dims[1])
    exports = {c: lat.halo_export() for c, lat in blocks.items()}
    for (cx, cy), lat in blocks.items():
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                if (dx == 0 and dy == 0) or (dx and not lat.ghost[0]) or (dy and not lat.ghost[1]):
                    continue
                lat.halo_connect((dx + 1) * 3 + dy + 1, exports[((cx + dx) % xs, (cy + dy) % ys)])
        lat.halo_finalize()


def run_blocks(blocks, n_steps, omega=None, chunk=None):
    """Advance every lattice of a `connect_blocks` topology by n_steps, safely: all blocks take the same launch groups
    in lockstep, and when two blocks share a device each group is drained before the next one is queued (see
    connect_blocks). `chunk` = steps per group (default: 1 on a shared device, all of them otherwise)."""
    lats = list(blocks.values()) if isinstance(blocks, dict) else list(blocks)
    shared = len({lat.device for lat in lats}) < len(lats)
    step = int(chunk) if chunk else (1 if shared else int(n_steps))
    done = 0
    while done < n_steps:
        n = min(step, n_steps - done)
        for lat in lats:
            lat.run(n, omega)
        if shared:
            for lat in lats:
                lat.sync()
        done += n


class _HistorySlot:
    """One device history slot and the handles parked in it. The slot returns to the free list when the last of its
    handles has been read or dropped."""

    def __init__(self, lattice, index, handles):
        self.lattice, self.index = lattice, index
        self.handles = [weakref.ref(h) for h in handles]
        for h in handles:
            h._hist = self
        lattice._parked.add(self)

    def read(self, which):
        L = self.lattice
        if not L._ctx:
            raise RuntimeError('history slot of a closed lattice (internal error)')
        out = np.empty(_SHAPES[which](L.nx, L.ny))
        N.check(L.lib.lbm_history_read(L._ctx, self.index, N.dptr(out) if which == 'rho' else None,
                                       N.dptr(out) if which == 'u' else None))
        return out

    def drain(self):
        for r in self.handles:
            h = r()
            if h is not None and h._hist is self:
                h.materialize()

    def __del__(self):
        try:
            if self.lattice._ctx and self.lattice._hist_free is not None:
                self.lattice._hist_free.append(self.index)
        except Exception:
            pass


_SHAPES = {'f': lambda nx, ny: (nx, ny, 9), 'rho': lambda nx, ny: (nx, ny), 'u': lambda nx, ny: (nx, ny, 2)}


class LatticeArray(np.lib.mixins.NDArrayOperatorsMixin):
    """ndarray-like view of one field of a `Lattice` at one time. Any numpy operation materialises it (one
    device->host copy, cached) — except scalar cell reads and whole-field np.amin / np.amax, which are answered from
    the device. `np.asarray(handle)` is the cached, READ-ONLY host copy (take `.copy()` to edit it); in-place edits
    through the handle itself (`h[i] = v`, `h += x`, `np.add(h, x, out=h)`) work as on the reference's arrays: the
    handle switches to a private writable copy and stops counting as the device's current state, so feeding it back
    into `lattice_boltzmann_step` uploads it."""

    __slots__ = ('_lattice', '_t', '_which', '_generation', '_value', '_hist', '_dirty', 'shape', '__weakref__')
    __array_priority__ = 100

    def __init__(self, lattice, t, which):
        self._lattice, self._t, self._which = lattice, t, which
        self._generation = lattice._generation
        self._value = None
        self._hist = None             # _HistorySlot while the field is parked on the device
        self._dirty = False           # written through __setitem__: the device copy no longer matches
        self.shape = lattice._shapes[which]

    dtype = np.dtype(np.float64)

    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape)))

    def _bring_current(self):
        L = self._lattice
        if self._value is None:
            if self._generation == L._generation and L.time < self._t <= L.api_time:
                L.flush(upto=self._t)
            if self._t != L.time or self._generation != L._generation:
                raise RuntimeError('stale LatticeArray: the lattice advanced without this handle being preserved '
                                   '(internal error)')

    @property
    def _on_device(self):
        """Not on the host and not parked: the field is (or will be) the lattice's current state."""
        return self._value is None and self._hist is None

    def materialize(self):
        if self._value is None and self._hist is not None:
            self._value = self._hist.read(self._which)
            self._value.setflags(write=False)
            self._hist = None
        if self._value is None:
            self._bring_current()
            L = self._lattice
            f, rho, u = L.fields(self._which == 'f', self._which == 'rho', self._which == 'u')
            self._value = {'f': f, 'rho': rho, 'u': u}[self._which]
            self._value.setflags(write=False)   # in-place edits must go through __setitem__ (tracked)
        return self._value

    # numpy protocols ----------------------------------------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        a = self.materialize()
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def _writable(self):
        """The host copy, made writable and marked as diverged from the device (in-place edits: f += x, out=f)."""
        a = self.materialize()
        if not a.flags.writeable:
            a = self._value = a.copy()
        self._dirty = True
        return a

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        inputs = tuple(x.materialize() if isinstance(x, LatticeArray) else x for x in inputs)
        outs = kwargs.get('out')
        if outs is not None:
            # the reference's results are ordinary writable ndarrays: `f += x` / np.add(f, x, out=f) must work on a
            # handle too. The handle then owns a private host copy and no longer counts as the device's current state.
            kwargs['out'] = tuple(x._writable() if isinstance(x, LatticeArray) else x for x in outs)
        res = getattr(ufunc, method)(*inputs, **kwargs)
        if outs is not None and any(isinstance(x, LatticeArray) for x in outs):
            if isinstance(res, tuple):
                return tuple(o if isinstance(o, LatticeArray) else r for o, r in zip(outs, res))
            return outs[0] if isinstance(outs[0], LatticeArray) else res
        return res

    def __array_function__(self, func, types, args, kwargs):
        if func in (np.amin, np.amax, np.min, np.max) and len(args) == 1 and not kwargs and args[0] is self \
                and self._on_device and self._which in ('rho', 'u'):
            return self._extremum(func in (np.amax, np.max))

        def conv(x):
            if isinstance(x, LatticeArray):
                return x.materialize()
            if isinstance(x, (list, tuple)):
                return type(x)(conv(v) for v in x)
            return x
        return func(*conv(args), **{k: conv(v) for k, v in kwargs.items()})

    def _extremum(self, want_max):
        self._bring_current()
        mn_r, mx_r, mn_u, mx_u = self._lattice.minmax()
        if self._which == 'rho':
            return np.float64(mx_r if want_max else mn_r)
        return np.float64(mx_u if want_max else mn_u)

    def min(self, *a, **k):
        return self._extremum(False) if not a and not k and self._on_device and self._which != 'f' else \
            self.materialize().min(*a, **k)

    def max(self, *a, **k):
        return self._extremum(True) if not a and not k and self._on_device and self._which != 'f' else \
            self.materialize().max(*a, **k)

    def __getitem__(self, index):
        if self._value is None and self._hist is None and type(index) is tuple and len(index) >= 2:
            ix, iy = index[0], index[1]
            if isinstance(ix, (int, np.integer)) and isinstance(iy, (int, np.integer)):
                # velocity[px, py, ...] every step (experiments.py:703-704): fetch one cell, not the lattice
                L = self._lattice
                if self._t != L.time or self._generation != L._generation:
                    self._bring_current()
                if not (-L.nx <= ix < L.nx and -L.ny <= iy < L.ny):
                    raise IndexError(f'index ({ix}, {iy}) is out of bounds for a lattice of shape ({L.nx}, {L.ny})')
                x, y = int(ix) % L.nx, int(iy) % L.ny
                cell = L._probe_sample(x, y, self._t) if self._which == 'u' else None
                if cell is None:
                    f, rho, u = L.fields(self._which == 'f', self._which == 'rho', self._which == 'u', (x, x + 1, y, y + 1))
                    cell = {'f': f, 'rho': rho, 'u': u}[self._which][0, 0]
                rest = tuple(i for i in index[2:] if i is not Ellipsis)
                return cell[rest] if rest else (cell if self._which != 'rho' else np.float64(cell))
        return self.materialize()[index]

    def __setitem__(self, index, value):
        self._writable()[index] = value

    def __len__(self):
        return self.shape[0]

    def __iter__(self):
        return iter(self.materialize())

    def __getattr__(self, name):
        # anything else an ndarray offers (copy, sum, mean, T, astype, tobytes, ...)
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __repr__(self):
        state = 'host' if self._value is not None else 'parked' if self._hist is not None else 'device'
        return f'LatticeArray({self._which}, t={self._t}, shape={self.shape}, {state})'
