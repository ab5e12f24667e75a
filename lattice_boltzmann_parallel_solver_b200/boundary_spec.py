"""Boundary conditions as data.

The reference expresses a scenario's boundary as nested Python closures that overwrite entries of
`f_post_streaming` one after another (src/boundary_utils.py:52-53, 103-106, 164-203). The fused kernel cannot
call Python per cell, so every closure factory of `boundary_conditions.py` returns a `BoundaryOp` that carries a
declarative description of the same overwrite, and `compile_ops` folds an ORDERED list of them into what
include/lbm_b200.h consumes: a one-byte kind per cell plus a small table of kinds (per population: pull /
bounce-back [- K] / constant / outlet copy). Later ops overwrite earlier ones slot by slot, exactly as the
closures' assignments do.

This is set-up code (runs once per scenario, O(boundary cells)); nothing here is on the time-step path.
"""
import numpy as np

from . import _native as N

OPP = (0, 3, 4, 1, 2, 7, 8, 5, 6)                 # src/lattice_boltzmann_method.py:37-39
CX = (0, 1, 0, -1, 0, 1, -1, -1, 1)               # src/lattice_boltzmann_method.py:14-26
CY = (0, 0, 1, 0, -1, 1, 1, -1, -1)
W = (4 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 36, 1 / 36, 1 / 36, 1 / 36)   # :50-52


def rule(kind, row=0):
    return kind | (row << 3)


class KindTable:
    """Interned cell kinds + the K / constant rows they refer to."""

    def __init__(self):
        self.kinds = [(tuple([0] * 9), 0, 0)]      # (rules, flags, skip_store); kind 0 = fluid
        self.index = {self.kinds[0]: 0}
        self.k_rows = [tuple([0.0] * 9)]
        self.c_rows = []
        self.rho_in = 0.0
        self.rho_out = 0.0

    def intern(self, desc):
        i = self.index.get(desc)
        if i is None:
            i = len(self.kinds)
            if i >= 256:
                raise ValueError('more than 256 distinct cell kinds')
            self.kinds.append(desc)
            self.index[desc] = i
        return i

    def _row(self, table, values, limit):
        # bit patterns decide identity (-0.0 and 0.0 are different constants)
        key = tuple(float(v) for v in values)
        for j, r in enumerate(table):
            if np.array_equal(np.array(r).view(np.uint64), np.array(key).view(np.uint64)):
                return j
        if len(table) >= limit:
            raise ValueError('more than %d constant rows' % limit)
        table.append(key)
        return len(table) - 1

    def k_row(self, values):
        return self._row(self.k_rows, values, 32)

    def c_row(self, values):
        return self._row(self.c_rows, values, 32)


class KindMap:
    """kind byte per cell of the (nx, ny) array the reference would hand to the closures."""

    def __init__(self, shape, table=None):
        self.shape = tuple(int(s) for s in shape)
        self.map = np.zeros(self.shape, dtype=np.uint8)
        self.table = table or KindTable()

    def modify(self, index, fn):
        """index: anything that indexes a (nx, ny) array (bool mask, slices, integer arrays).
        fn(rules: list[9], flags, skip) -> (rules, flags, skip) for one kind."""
        cur = self.map[index]
        if np.size(cur) == 0:
            return
        trans = {}
        for k in np.unique(cur):
            rules, flags, skip = self.table.kinds[int(k)]
            r2, f2, s2 = fn(list(rules), flags, skip)
            trans[int(k)] = self.table.intern((tuple(int(v) for v in r2), int(f2), int(s2)))
        lut = np.arange(256, dtype=np.uint8)
        for a, b in trans.items():
            lut[a] = b
        self.map[index] = lut[cur]

    def view(self, sx, sy):
        return _KindView(self, sx, sy)

    # ---- the overwrites of the reference's closures ------------------------------------------------------
    def bounce(self, index, dirs, k_values=None):
        """f_post[index, opp(d)] = f_pre[index, d] - K_d  (boundary_conditions.py:108-109, 207-210)."""
        row = self.table.k_row(k_values) if k_values is not None else 0

        def fn(rules, flags, skip):
            for d in dirs:
                rules[OPP[d]] = rule(N.RULE_BOUNCE, row)
            return rules, flags, skip
        self.modify(index, fn)

    def constant(self, index, values):
        """f_post[index, i] = values[i] for all nine i  (boundary_conditions.py:250-251)."""
        row = self.table.c_row(values)

        def fn(rules, flags, skip):
            return [rule(N.RULE_CONST, row)] * 9, flags, skip
        self.modify(index, fn)

    def flag(self, index, bits=0, skip_bits=0, rules_for=None):
        def fn(rules, flags, skip):
            if rules_for:
                for i, r in rules_for.items():
                    rules[i] = r
            return rules, flags | bits, skip | skip_bits
        self.modify(index, fn)

    def to_desc(self):
        """-> (BcDesc, keepalive) for lbm_create / lbm_bc_apply."""
        t = self.table
        kinds = np.zeros(len(t.kinds), dtype=N.KIND_DTYPE)
        for i, (rules, flags, skip) in enumerate(t.kinds):
            kinds[i]['rule'] = rules
            kinds[i]['flags'] = flags
            kinds[i]['skip_store'] = skip
        ktab = np.ascontiguousarray(np.array(t.k_rows, dtype=np.float64).reshape(-1, 9))
        ctab = np.ascontiguousarray(np.array(t.c_rows, dtype=np.float64).reshape(-1, 9)) if t.c_rows else \
            np.zeros((0, 9))
        kmap = np.ascontiguousarray(self.map)
        d = N.BcDesc()
        d.n_kinds = len(kinds)
        d.kinds = kinds.ctypes.data_as(N.C.POINTER(N.Kind))
        d.n_k_rows = len(ktab)
        d.k_table = N.dptr(ktab)
        d.n_c_rows = len(ctab)
        d.c_table = N.dptr(ctab) if len(ctab) else None
        d.pbc_rho_in = float(t.rho_in)
        d.pbc_rho_out = float(t.rho_out)
        d.kind_map = kmap.ctypes.data_as(N.C.POINTER(N.C.c_uint8))
        return d, (kinds, ktab, ctab, kmap)

    @property
    def is_trivial(self):
        return not self.map.any()

    def digest(self):
        """Content hash of the compiled description (kind bytes + kind table + K / constant rows): two bundles that
        compile to the same bytes are the same scenario for the device (lattice cache key)."""
        import hashlib
        t = self.table
        h = hashlib.sha256()
        h.update(repr(self.shape).encode())
        h.update(np.ascontiguousarray(self.map).tobytes())
        h.update(repr(t.kinds).encode())
        h.update(np.array(t.k_rows, dtype=np.float64).tobytes())
        h.update(np.array(t.c_rows, dtype=np.float64).tobytes())
        h.update(np.array([t.rho_in, t.rho_out], dtype=np.float64).tobytes())
        return h.hexdigest()


class _KindView:
    """A rectangular window of a KindMap (the reference applies inlet/outlet to f_post[1:-1, 1:-1],
    src/boundary_utils.py:166-173)."""

    def __init__(self, parent, sx, sy):
        self.parent, self.sx, self.sy = parent, sx, sy
        self.shape = parent.map[sx, sy].shape
        self.table = parent.table

    def _abs(self, index):
        nx, ny = self.parent.shape
        gx = np.arange(nx)[self.sx]
        gy = np.arange(ny)[self.sy]
        sel = np.zeros(self.shape, dtype=bool)
        sel[index] = True
        full = np.zeros((nx, ny), dtype=bool)
        full[np.ix_(gx, gy)] = sel
        return full

    def modify(self, index, fn):
        self.parent.modify(self._abs(index), fn)

    bounce = KindMap.bounce
    constant = KindMap.constant
    flag = KindMap.flag


def compile_ops(shape, ops):
    """Fold an ordered list of (BoundaryOp, window) into a KindMap. window is None (whole array) or a pair of
    slices."""
    km = KindMap(shape)
    for op, window in ops:
        target = km if window is None else km.view(*window)
        op.emit(target)
    return km
