"""A CPU stand-in for liblbm_b200.so, for TESTS of the host-side handle logic only (engine.py, the lazy results of
lattice_boltzmann_step). It implements the subset of include/lbm_b200.h that a periodic, boundary-free lattice
needs, with the oracle as its arithmetic. It is never importable from the product package."""
import ctypes as C

import numpy as np

from oracle import lbm_numpy as onp


def _arr(ptr, shape):
    return np.ctypeslib.as_array(ptr, shape=shape) if ptr else None


class _Ctx:
    def __init__(self, nx, ny):
        self.nx, self.ny = nx, ny
        self.state = None        # (f, rho, u) of the current time, reference semantics
        self.t = 0
        self.launches = 0
        self.steps_log = []      # omegas, for assertions
        self.calls = []          # ('step', n) / ('fields', 1) / ('probe_config', x, y) / ('probe_read', t0, n) in call order
        self.probe = None        # (x, y, capacity, t_config)
        self.samples = {}        # t -> (ux, uy)


class FakeLib:
    def __init__(self):
        self.ctxs = {}
        self.next = 1
        self.err = b''

    # -- context -------------------------------------------------------------------------------------------
    def lbm_device_count(self):
        return 1

    def lbm_last_error(self):
        return self.err

    def lbm_create(self, device, nx, ny, gx, gy, bc, out):
        assert not bc, 'the fake only knows boundary-free lattices'
        h = self.next
        self.next += 1
        self.ctxs[h] = _Ctx(nx, ny)
        out._obj.value = h
        return 0

    def _c(self, ctx):
        return self.ctxs[ctx.value if hasattr(ctx, 'value') else ctx]

    def lbm_destroy(self, ctx):
        self.ctxs.pop(ctx.value, None)
        return 0

    def lbm_set_bc_mode(self, ctx, mode):
        return 0

    def lbm_set_option(self, ctx, name, value):
        return 0

    def lbm_stream(self, ctx):
        return 0

    def lbm_device_bytes(self, ctx):
        return 0

    def lbm_launch_count(self, ctx):
        return self._c(ctx).launches

    def lbm_sync(self, ctx):
        return 0

    def lbm_upload(self, ctx, f, rho, u, omega):
        c = self._c(ctx)
        c.state = (np.array(_arr(f, (c.nx, c.ny, 9))), np.array(_arr(rho, (c.nx, c.ny))), np.array(_arr(u, (c.nx, c.ny, 2))))
        c.t = 0
        c.samples = {}
        return 0

    def lbm_step(self, ctx, omega, n):
        c = self._c(ctx)
        if c.state is None:
            self.err = b'lbm_step before lbm_upload'
            return 3
        c.calls.append(('step', n))
        for _ in range(n):
            c.state = onp.step(*c.state, omega)
            c.t += 1
            c.launches += 1
            c.steps_log.append(omega)
            if c.probe is not None:
                c.samples[c.t] = np.array(c.state[2][c.probe[0], c.probe[1]])
        return 0

    def lbm_materialize_region(self, ctx, x0, x1, y0, y1, f, rho, u):
        c = self._c(ctx)
        if c.t == 0:
            self.err = b'no step taken since the state was loaded'
            return 3
        c.launches += 1
        c.calls.append(('fields', 1))
        for ptr, src, tail in ((f, c.state[0], (9,)), (rho, c.state[1], ()), (u, c.state[2], (2,))):
            if ptr:
                _arr(ptr, (x1 - x0, y1 - y0) + tail)[...] = src[x0:x1, y0:y1]
        return 0

    def lbm_probe_config(self, ctx, x, y, capacity):
        c = self._c(ctx)
        c.probe, c.samples = (x, y, capacity, c.t), {}
        c.calls.append(('probe_config', x, y))
        return 0

    def lbm_probe_read(self, ctx, t0, n, out):
        c = self._c(ctx)
        if c.probe is None or t0 < 1 or t0 + n - 1 > c.t or c.t - t0 >= c.probe[2] or any(t not in c.samples for t in range(t0, t0 + n)):
            self.err = b'lbm_probe_read: steps not in the ring'
            return 1
        c.calls.append(('probe_read', t0, n))
        _arr(out, (n, 2))[...] = [c.samples[t] for t in range(t0, t0 + n)]
        return 0

    def lbm_minmax(self, ctx, x0, x1, y0, y1, out):
        c = self._c(ctx)
        r, u = c.state[1][x0:x1, y0:y1], c.state[2][x0:x1, y0:y1]
        _arr(out, (4,))[...] = [r.min(), r.max(), u.min(), u.max()]
        return 0
