"""A CPU stand-in for liblbm_b200.so, for TESTS of the host-side logic only (engine.py, the lazy results of
lattice_boltzmann_step, the reference's own drivers running on the drop-in modules without a GPU). It implements the
subset of include/lbm_b200.h those tests reach — stateless operators, contexts with the kind-rule boundary description
(pull / bounce [- K] / constant / outlet, the pressure-periodic source rows) and a self-periodic ghost ring — with the oracle
as its arithmetic. It is never importable from the product package."""
import ctypes as C

import numpy as np

from oracle import lbm_numpy as onp


def _arr(ptr, shape):
    return np.ctypeslib.as_array(ptr, shape=shape) if ptr else None


OPP = (0, 3, 4, 1, 2, 7, 8, 5, 6)


class _Ctx:
    def __init__(self, nx, ny, ghost=(0, 0), bc=None):
        self.nx, self.ny = nx, ny
        self.ghost = ghost
        self.bc = bc             # None or dict(kind_map, kinds, ktab, ctab)
        self.state = None        # (f, rho, u) of the current time, reference semantics
        self.t = 0
        self.launches = 0
        self.steps_log = []      # omegas, for assertions
        self.calls = []          # ('step', n) / ('fields', 1) / ('probe_config', x, y) / ('probe_read', t0, n) in call order
        self.probe = None        # (x, y, capacity, t_config)
        self.samples = {}        # t -> (ux, uy)
        self.hist = []           # history slots: None or (rho, u)


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

    @staticmethod
    def _parse(bc, nx, ny):
        d = bc._obj if hasattr(bc, '_obj') else bc.contents
        kinds = [(tuple(d.kinds[k].rule), int(d.kinds[k].flags), int(d.kinds[k].skip_store)) for k in range(d.n_kinds)]
        return {'kind_map': np.array(np.ctypeslib.as_array(d.kind_map, shape=(nx, ny))),
                'kinds': kinds,
                'ktab': np.array(np.ctypeslib.as_array(d.k_table, shape=(d.n_k_rows, 9))),
                'ctab': np.array(np.ctypeslib.as_array(d.c_table, shape=(d.n_c_rows, 9))) if d.n_c_rows else np.zeros((0, 9)),
                # pressure-periodic boundary (flags 2 | 4 mark its source rows; skip_store is a device detail)
                'pbc': (float(d.pbc_rho_in), float(d.pbc_rho_out)) if any(fl & 6 for _, fl, _ in kinds) else None}

    @staticmethod
    def _apply_pbc(rho_in, rho_out, f_pre, rho, u):
        """periodic_with_pressure_variations, x variant, in place on f_pre (src/boundary_conditions.py:337-344)."""
        ly = f_pre.shape[1]
        feq_m2, feq_p1 = onp.equilibrium(rho[-2], u[-2]), onp.equilibrium(rho[1], u[1])
        feq_in, feq_out = onp.equilibrium(np.ones(ly) * rho_in, u[-2]), onp.equilibrium(np.ones(ly) * rho_out, u[1])
        for d in (1, 5, 8):
            f_pre[0, :, d] = feq_in[:, d] + (f_pre[-2, :, d] - feq_m2[:, d])
        for d in (3, 6, 7):
            f_pre[-1, :, d] = feq_out[:, d] + (f_pre[1, :, d] - feq_p1[:, d])
        return f_pre

    @staticmethod
    def _apply_rules(desc, f_pre, f_post, f_prev):
        """The closures' overwrites as data (include/lbm_b200.h): bounce-back (- wall term), inlet constants, outlet copy."""
        km = desc['kind_map']
        for k, (rules, flags, skip) in enumerate(desc['kinds']):
            if k == 0:
                continue
            cells = km == k
            if not cells.any():
                continue
            for i, r in enumerate(rules):
                typ, row = r & 7, r >> 3
                if typ == 1:      # bounce: f_post[i] = f_pre[opp i] - K[row][opp i]
                    f_post[cells, i] = np.subtract(f_pre[cells, OPP[i]], desc['ktab'][row][OPP[i]]) if row else f_pre[cells, OPP[i]]
                elif typ == 2:    # inlet constants
                    f_post[cells, i] = desc['ctab'][row][i]
                elif typ == 3:    # outlet: the step's input f of the row before
                    xs, ys = np.nonzero(cells)
                    f_post[xs, ys, i] = f_prev[xs - 1, ys, i]
        return f_post

    def lbm_bc_apply(self, device, nx, ny, bc, f_pre, f_post, f_prev):
        desc = self._parse(bc, nx, ny)
        self._apply_rules(desc, _arr(f_pre, (nx, ny, 9)), _arr(f_post, (nx, ny, 9)), _arr(f_prev, (nx, ny, 9)) if f_prev else None)
        return 0

    def lbm_pbc_apply(self, device, nx, ny, rho_in, rho_out, rho, u, f_pre):
        self._apply_pbc(rho_in, rho_out, _arr(f_pre, (nx, ny, 9)), _arr(rho, (nx, ny)), _arr(u, (nx, ny, 2)))
        return 0

    def lbm_create(self, device, nx, ny, gx, gy, bc, out):
        desc = self._parse(bc, nx, ny) if bc else None
        h = self.next
        self.next += 1
        self.ctxs[h] = _Ctx(nx, ny, (gx, gy), desc)
        out._obj.value = h
        return 0

    # -- halo: one rank, every neighbour is the lattice itself (communication() with a 1x1 topology) ------------------
    def lbm_halo_export_handle(self, ctx, out):
        return 0

    def lbm_halo_connect(self, ctx, slot, peer):
        return 0

    def lbm_halo_finalize(self, ctx):
        return 0

    # -- stateless operators -------------------------------------------------------------------------------------
    def lbm_equilibrium(self, device, n, rho, u, out):
        _arr(out, (n, 9))[...] = onp.equilibrium(_arr(rho, (n,)), _arr(u, (n, 2)))
        return 0

    def lbm_density(self, device, n, f, out):
        _arr(out, (n,))[...] = onp.density(_arr(f, (n, 9)))
        return 0

    def lbm_velocity(self, device, n, rho, f, out):
        _arr(out, (n, 2))[...] = onp.velocity(_arr(rho, (n,)), _arr(f, (n, 9)))
        return 0

    def lbm_streaming(self, device, nx, ny, f, out):
        _arr(out, (nx, ny, 9))[...] = onp.stream(_arr(f, (nx, ny, 9)))
        return 0

    def _one_step(self, c, omega):
        """collide with the given moments -> ghost exchange -> stream -> kind rules -> moments
        (src/lattice_boltzmann_method.py:213-226 with the closures' overwrites as data, include/lbm_b200.h)."""
        f, rho, u = c.state
        if c.bc is None and c.ghost == (0, 0):
            return onp.step(f, rho, u, omega)
        f_pre = onp.collide(f, rho, u, omega)
        if c.ghost != (0, 0):
            assert c.ghost == (1, 1)
            f_pre = onp.self_exchange(f_pre)
        if c.bc is not None and c.bc['pbc']:
            f_pre = self._apply_pbc(*c.bc['pbc'], f_pre, rho, u)
        f_post = onp.stream(f_pre)
        if c.bc is not None:
            f_post = self._apply_rules(c.bc, f_pre, f_post, f)
        rho2 = onp.density(f_post)
        return f_post, rho2, onp.velocity(rho2, f_post)

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
            c.state = self._one_step(c, omega)
            c.t += 1
            c.launches += 1
            c.steps_log.append(omega)
            if c.probe is not None:
                c.samples[c.t] = np.array(c.state[2][c.probe[0], c.probe[1]])
        return 0

    def lbm_run_host(self, ctx, f, rho, u, omega, n, fo, ro, uo):
        c = self._c(ctx)
        rc = self.lbm_upload(ctx, f, rho, u, omega) or self.lbm_step(ctx, omega, n)
        if rc:
            return rc
        c.calls.append(('run_host', n))
        for ptr, src, tail in ((fo, c.state[0], (9,)), (ro, c.state[1], ()), (uo, c.state[2], (2,))):
            if ptr:
                _arr(ptr, (c.nx, c.ny) + tail)[...] = src
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

    def lbm_history_config(self, ctx, n_slots):
        c = self._c(ctx)
        if c.ghost[0] >= 2:
            self.err = b'lbm_history_config: slabs keep no history'
            return 1
        c.hist = [None] * n_slots
        c.calls.append(('hist_config', n_slots))
        return 0

    def lbm_history_store(self, ctx, slot):
        c = self._c(ctx)
        if c.t == 0 or not 0 <= slot < len(c.hist):
            self.err = b'lbm_history_store: bad slot or no step taken'
            return 1
        c.hist[slot] = (c.state[1].copy(), c.state[2].copy())
        c.launches += 1
        c.calls.append(('hist_store', slot))
        return 0

    def lbm_history_read(self, ctx, slot, rho, u):
        c = self._c(ctx)
        if not 0 <= slot < len(c.hist) or c.hist[slot] is None:
            self.err = b'lbm_history_read: bad slot'
            return 1
        c.calls.append(('hist_read', slot))
        if rho:
            _arr(rho, (c.nx, c.ny))[...] = c.hist[slot][0]
        if u:
            _arr(u, (c.nx, c.ny, 2))[...] = c.hist[slot][1]
        return 0

    def lbm_minmax(self, ctx, x0, x1, y0, y1, out):
        c = self._c(ctx)
        r, u = c.state[1][x0:x1, y0:y1], c.state[2][x0:x1, y0:y1]
        _arr(out, (4,))[...] = [r.min(), r.max(), u.min(), u.max()]
        return 0
