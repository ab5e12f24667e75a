"""CPU ORACLE (numpy) for the D2Q9 fp64 time step — TEST INFRASTRUCTURE, NOT PRODUCT.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module. The product path (`lattice_boltzmann_parallel_solver_b200/`) never does and fails loudly without
its CUDA library.

This is a restatement, written from the arithmetic contract in SURVEY.md §8(a), of the reference's
numpy algorithm; every function cites the reference file:line it follows (paths relative to
/root/reference). Parity status: PINNED — `tests/test_oracle_golden.py` checks it bit-for-bit (sha256 of
the float64 bytes) against fixtures produced by running the unmodified reference
(`tests/golden/make_goldens.py`), including the reference's own 12-sample golden
`tests/von_karman_vortex_shedding/vel_at_p.npy` and the first 2001 samples of the cluster trace
`figures/von_karman_vortex_shedding/reynold_strouhal/vel_at_p_100.npy`.

Layout as in the reference: f[x, y, i] (C-contiguous, i fastest), rho[x, y], u[x, y, 2], all float64.
Every `fl(.)` of the contract is one numpy ufunc call below, so the association order is explicit.
"""
import numpy as np

# D2Q9 lattice constants — src/lattice_boltzmann_method.py:14-26 (c_i), :37-39 (opposite), :50-52 (w_i)
CX = (0, 1, 0, -1, 0, 1, -1, -1, 1)
CY = (0, 0, 1, 0, -1, 1, 1, -1, -1)
OPP = (0, 3, 4, 1, 2, 7, 8, 5, 6)
W = (4 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 36, 1 / 36, 1 / 36, 1 / 36)
# c_s = 1/np.sqrt(3); c_s**2 = 0x1.5555555555557p-2, two ulp above 1/3 — src/boundary_conditions.py:186,304
CS2 = float((1 / np.sqrt(3)) ** 2)


# ---------------------------------------------------------------------------------------------------------
# a2-a5: moments, equilibrium, streaming
# ---------------------------------------------------------------------------------------------------------
def equilibrium(rho, u):
    """src/lattice_boltzmann_method.py:162-188.

    cu_i = fl(cx*ux + cy*uy) (a single rounding: the products by 0/+-1 are exact, :179);
    uu = fl(fl(sqrt(fl(fl(ux^2)+fl(uy^2))))^2) — norm first, then square (:185);
    feq_i = fl(fl(w_i*rho) * fl(fl(fl(1+fl(3cu)) + fl(4.5*fl(cu^2))) - fl(1.5*uu))) (:181-186).
    """
    rho = np.asarray(rho, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    assert rho.shape == u.shape[:-1]
    ux, uy = u[..., 0], u[..., 1]
    nrm = np.sqrt(ux * ux + uy * uy)
    t = 1.5 * (nrm * nrm)
    out = np.empty(rho.shape + (9,), dtype=np.float64)
    diag_p = ux + uy      # c = (1, 1); c = (-1,-1) is its exact negation
    diag_m = uy - ux      # c = (-1, 1); c = (1,-1) is its exact negation
    cu = (None, ux, uy, -ux, -uy, diag_p, diag_m, -diag_p, -diag_m)
    for i in range(9):
        wr = W[i] * rho
        if i == 0:
            # cu = 0: fl(fl(fl(1+0)+0) - t) = fl(1 - t)
            out[..., 0] = wr * (1.0 - t)
        else:
            c = cu[i]
            out[..., i] = wr * (((1.0 + 3.0 * c) + 4.5 * (c * c)) - t)
    return out


def density(f):
    """src/lattice_boltzmann_method.py:93-105. np.sum over 9 contiguous addends = numpy's pairwise-8
    kernel: ((f0+f1)+(f2+f3)) + ((f4+f5)+(f6+f7)), then + f8."""
    assert f.shape[-1] == 9
    return (((f[..., 0] + f[..., 1]) + (f[..., 2] + f[..., 3])) +
            ((f[..., 4] + f[..., 5]) + (f[..., 6] + f[..., 7]))) + f[..., 8]


def velocity(rho, f):
    """src/lattice_boltzmann_method.py:108-137: ((f1+f5)+f8) - ((f3+f6)+f7) over rho, and the y analogue;
    0 where rho == 0 (:126,:132)."""
    assert f.shape[-1] == 9
    jx = ((f[..., 1] + f[..., 5]) + f[..., 8]) - ((f[..., 3] + f[..., 6]) + f[..., 7])
    jy = ((f[..., 2] + f[..., 5]) + f[..., 6]) - ((f[..., 4] + f[..., 7]) + f[..., 8])
    u = np.zeros(rho.shape + (2,), dtype=np.float64)
    nz = rho != 0
    np.divide(jx, rho, out=u[..., 0], where=nz)
    np.divide(jy, rho, out=u[..., 1], where=nz)
    return u


def stream(f):
    """src/lattice_boltzmann_method.py:140-159: out[x,y,i] = in[(x-cx_i) mod lx, (y-cy_i) mod ly, i]
    (pull form of the 9 np.roll calls), periodic over whatever array is given, ghost rows included."""
    assert f.shape[-1] == 9
    lx, ly = f.shape[0], f.shape[1]
    out = np.empty_like(f)
    for i in range(9):
        # destination rows [x_dst] take source rows [x_src]; one wrapped row/column per non-zero component
        for xd, xsrc in _shift_slices(lx, CX[i]):
            for yd, ysrc in _shift_slices(ly, CY[i]):
                out[xd, yd, i] = f[xsrc, ysrc, i]
    return out


def _shift_slices(n, c):
    """(dst, src) slice pairs realising dst = (src + c) mod n along one axis."""
    if c == 0:
        return ((slice(None), slice(None)),)
    if c == 1:
        return ((slice(1, n), slice(0, n - 1)), (slice(0, 1), slice(n - 1, n)))
    return ((slice(0, n - 1), slice(1, n)), (slice(n - 1, n), slice(0, 1)))


def collide(f, rho, u, omega):
    """src/lattice_boltzmann_method.py:213-215: f + (feq(rho,u) - f)*omega with the CALLER's moments."""
    feq = equilibrium(rho, u)
    return f + (feq - f) * omega


# ---------------------------------------------------------------------------------------------------------
# a6-a11: boundary operators (closures mutate and return their array argument, as the reference's do)
# ---------------------------------------------------------------------------------------------------------
def wall_directions(mask):
    """src/boundary_conditions.py:8-28: the first fully-set edge decides which three populations bounce."""
    if np.all(mask[0, :]):
        return (1, 5, 8)
    if np.all(mask[-1, :]):
        return (3, 6, 7)
    if np.all(mask[:, 0]):
        return (4, 7, 8)
    if np.all(mask[:, -1]):
        return (2, 5, 6)
    return ()


def rigid_wall(mask):
    """src/boundary_conditions.py:78-113: f_post[mask, opp(d)] = f_pre[mask, d]."""
    assert mask.dtype == bool
    dirs = wall_directions(mask)

    def bc(f_pre, f_post):
        for d in dirs:
            f_post[mask, OPP[d]] = f_pre[mask, d]
        return f_post
    return bc


def moving_wall_constants(dirs, u_w, avg_density):
    """K_d = ((2*w_d)*avg_density) * ((c_d . u_w)/c_s**2) — src/boundary_conditions.py:207-210."""
    k = {}
    for d in dirs:
        cdotu = CX[d] * u_w[0] + CY[d] * u_w[1]
        k[d] = 2 * W[d] * avg_density * np.divide(cdotu, CS2)
    return k


def moving_wall(mask, u_w, avg_density):
    """src/boundary_conditions.py:170-214: bounce-back minus the wall momentum term K_d."""
    assert mask.dtype == bool
    dirs = wall_directions(mask)
    k = moving_wall_constants(dirs, u_w, avg_density)

    def bc(f_pre, f_post):
        for d in dirs:
            f_post[mask, OPP[d]] = f_pre[mask, d] - k[d]
        return f_post
    return bc


def plate_cells(mask):
    """src/boundary_conditions.py:31-75 and :131-134: for a thin vertical plate marked on column x0 over
    [y_lo, y_hi] returns (x0, y_lo, y_hi). Full 3-population bounce on y_lo+1..y_hi-1 of columns x0
    (1,5,8) and x0+1 (3,6,7); the four corner cells bounce only two populations (:50-55)."""
    idx = np.argwhere(mask)
    x0 = int(idx[:, 0].min())
    assert int(idx[:, 0].max()) == x0, 'oracle restates the one-column plate the reference uses'
    return x0, int(idx[:, 1].min()), int(idx[:, 1].max())


def rigid_object(mask):
    """src/boundary_conditions.py:116-167 (the reference also clears the corner cells in the caller's mask,
    :73-75; reproduced)."""
    assert mask.dtype == bool
    x0, ylo, yhi = plate_cells(mask)
    mask[x0, ylo] = False
    mask[x0, yhi] = False
    lx = mask.shape[0]
    x1 = (x0 + 1) % lx

    def bc(f_pre, f_post):
        for d in (1, 5, 8):
            f_post[x0, ylo + 1:yhi, OPP[d]] = f_pre[x0, ylo + 1:yhi, d]
        for d in (3, 6, 7):
            f_post[x1, ylo + 1:yhi, OPP[d]] = f_pre[x1, ylo + 1:yhi, d]
        for (x, y, dirs) in ((x0, yhi, (1, 8)), (x0, ylo, (1, 5)), (x0 + 1, yhi, (3, 7)), (x0 + 1, ylo, (3, 6))):
            for d in dirs:
                f_post[x, y, OPP[d]] = f_pre[x, y, d]
        return f_post
    return bc


def inlet_constants(rho_in, u_in):
    """The 9 values src/boundary_conditions.py:232-237 precomputes: feq(rho_in*1, (u_in, 0))."""
    return equilibrium(np.ones((1, 1)) * rho_in, np.array([[[u_in, 0.0]]], dtype=np.float64))[0, 0]


def inlet(shape, rho_in, u_in):
    """src/boundary_conditions.py:217-254: column x=0 of f_post := feq(rho_in, (u_in,0)) for all 9."""
    c = inlet_constants(rho_in, u_in)

    def bc(f_post):
        f_post[0, :, :] = c
        return f_post
    return bc


def outlet():
    """src/boundary_conditions.py:257-283: f_post[-1,:,d] = f_previous[-2,:,d], d in (3,6,7)."""
    def bc(f_prev, f_post):
        for d in (3, 6, 7):
            f_post[-1, :, d] = f_prev[-2, :, d]
        return f_post
    return bc


def pbc_pressure_x(p_in, p_out):
    """src/boundary_conditions.py:286-348, x-direction case (:305-311, :337-344), in place on f_pre:
    f_pre[0,:,d]  = feq_d(rho_in,  u[-2,:]) + (f_pre[-2,:,d] - feq_d(rho,u)[-2,:]) for d in (1,5,8)
    f_pre[-1,:,d] = feq_d(rho_out, u[1,:])  + (f_pre[1,:,d]  - feq_d(rho,u)[1,:])  for d in (3,6,7)
    with rho_in = p_in/c_s**2, rho_out = p_out/c_s**2 (the non-1/3 c_s**2)."""
    rho_in = np.divide(p_in, CS2)
    rho_out = np.divide(p_out, CS2)

    def bc(f_pre, rho, u):
        ly = f_pre.shape[1]
        feq_m2 = equilibrium(rho[-2], u[-2])
        feq_p1 = equilibrium(rho[1], u[1])
        feq_in = equilibrium(np.ones(ly) * rho_in, u[-2])
        feq_out = equilibrium(np.ones(ly) * rho_out, u[1])
        for d in (1, 5, 8):
            f_pre[0, :, d] = feq_in[:, d] + (f_pre[-2, :, d] - feq_m2[:, d])
        for d in (3, 6, 7):
            f_pre[-1, :, d] = feq_out[:, d] + (f_pre[1, :, d] - feq_p1[:, d])
        return f_pre
    return bc


# ---------------------------------------------------------------------------------------------------------
# a12: scenario bundles (order matters) — src/boundary_utils.py
# ---------------------------------------------------------------------------------------------------------
def _edge(shape, which):
    m = np.zeros(shape, dtype=bool)
    if which == 'y0':
        m[:, 0] = True
    elif which == 'y1':
        m[:, -1] = True
    elif which == 'x0':
        m[0, :] = True
    else:
        m[-1, :] = True
    return m


def couette_bc(lx, ly, U, avg_density):
    """src/boundary_utils.py:9-55: rigid wall at y=ly-1 first, then the moving wall at y=0."""
    top = rigid_wall(_edge((lx, ly), 'y1'))
    bottom = moving_wall(_edge((lx, ly), 'y0'), np.array([U, 0]), avg_density)

    def boundary(f_pre, f_post, rho=None, u=None, f_prev=None):
        return bottom(f_pre, top(f_pre, f_post))
    return boundary


def poiseuille_bc(lx, ly, p_in, p_out):
    """src/boundary_utils.py:58-110: pressure BC on f_pre, STREAM AGAIN (:103-104), then both rigid walls
    over all x (virtual columns included) reading the modified f_pre (:105-106)."""
    pbc = pbc_pressure_x(p_in, p_out)
    bottom = rigid_wall(_edge((lx, ly), 'y0'))
    top = rigid_wall(_edge((lx, ly), 'y1'))

    def boundary(f_pre, f_post, rho, u, f_prev=None):
        f_pre = pbc(f_pre, rho, u)
        f_post = stream(f_pre)
        return top(f_pre, bottom(f_pre, f_post))
    return boundary


# a14 topology math — src/parallelization_utils.py:55-200
def xy_size(n):
    """:55-92 — most-square factorisation with x_size <= y_size; primes > 2 raise."""
    if n > 2 and all(n % i for i in range(2, n)):
        raise Exception('This implementation does not work if number of nodes is a prime (excluding 1 and 2)')
    if n <= 1:
        return 1, 1
    lo = hi = int(np.ceil(np.sqrt(n)))
    while lo * hi != n:
        if lo * hi > n:
            lo -= 1
        else:
            hi += 1
    return lo, hi


def block_origin(c, L, P):
    return c * (L // P)


def block_extent(c, L, P):
    """:95-120 — L//P per rank, the last rank takes the remainder."""
    return L - (L // P) * (P - 1) if c + 1 == P else L // P


def owns(c, g, L, P):
    """:165-200."""
    lo = c * (L // P)
    hi = (c + 1) * (L // P) - 1 if c != P - 1 else L - 1
    return lo <= g <= hi


def to_local(c, g, L, P):
    """:123-137 — global index -> index in the ghost-padded local array."""
    return int(g - c * (L // P)) + 1


def karman_parallel_bc(coord, nlx, nly, lx, ly, xs, ys, rho_in, u_in, plate):
    """src/boundary_utils.py:113-205 on ghost-padded local arrays (nlx+2, nly+2): inlet on the interior view
    if the rank owns x=0 (:166-168); outlet if it owns both lx-1 and lx-2, NotImplementedError if only one
    (:171-176); plate columns lx//4 and lx//4+1, full bounce for y in [ly//2-d//2+1, ly//2+d//2-2], two-
    population corners at ly//2+d//2-1 and ly//2-d//2 (:145-201)."""
    cx, cy = coord
    cin = inlet_constants(rho_in, u_in)
    y_full = [y for y in range(ly // 2 - plate // 2 + 1, ly // 2 + plate // 2 - 1) if owns(cy, y, ly, ys)]
    yl = np.array([to_local(cy, y, ly, ys) for y in y_full], dtype=np.int64)
    y_top, y_bot = ly // 2 + plate // 2 - 1, ly // 2 - plate // 2
    has_in = owns(cx, 0, lx, xs)
    has_o1, has_o2 = owns(cx, lx - 1, lx, xs), owns(cx, lx - 2, lx, xs)
    has_l, has_r = owns(cx, lx // 4, lx, xs), owns(cx, lx // 4 + 1, lx, xs)

    def boundary(f_pre, f_post, rho=None, u=None, f_prev=None):
        if has_in:
            f_post[1, 1:-1, :] = cin
        if has_o1 and has_o2:
            for d in (3, 6, 7):
                f_post[-2, 1:-1, d] = f_prev[-3, 1:-1, d]
        elif has_o1 or has_o2:
            raise NotImplementedError
        if has_l:
            x = to_local(cx, lx // 4, lx, xs)
            for a, b in ((3, 1), (7, 5), (6, 8)):
                f_post[x, yl, a] = f_pre[x, yl, b]
            if owns(cy, y_top, ly, ys):
                y = to_local(cy, y_top, ly, ys)
                f_post[x, y, 3], f_post[x, y, 6] = f_pre[x, y, 1], f_pre[x, y, 8]
            if owns(cy, y_bot, ly, ys):
                y = to_local(cy, y_bot, ly, ys)
                f_post[x, y, 3], f_post[x, y, 7] = f_pre[x, y, 1], f_pre[x, y, 5]
        if has_r:
            x = to_local(cx, lx // 4 + 1, lx, xs)
            for a, b in ((1, 3), (5, 7), (8, 6)):
                f_post[x, yl, a] = f_pre[x, yl, b]
            if owns(cy, y_top, ly, ys):
                y = to_local(cy, y_top, ly, ys)
                f_post[x, y, 1], f_post[x, y, 5] = f_pre[x, y, 3], f_pre[x, y, 7]
            if owns(cy, y_bot, ly, ys):
                y = to_local(cy, y_bot, ly, ys)
                f_post[x, y, 1], f_post[x, y, 8] = f_pre[x, y, 3], f_pre[x, y, 6]
        return f_post
    return boundary


def karman_serial_bc(lx, ly, rho_in, u_in, plate):
    """milestoneQuickFunctionCalls.py:304-312 — inlet, outlet, rigid_object on un-padded arrays (the recipe
    that wrote the reference's f_i.npy goldens)."""
    i_bc = inlet((lx, ly), rho_in, u_in)
    o_bc = outlet()
    m = np.zeros((lx, ly), dtype=bool)
    m[lx // 4, ly // 2 - plate // 2:ly // 2 + plate // 2] = True
    p_bc = rigid_object(m)

    def boundary(f_pre, f_post, rho=None, u=None, f_prev=None):
        return p_bc(f_pre, o_bc(f_prev, i_bc(f_post)))
    return boundary


# ---------------------------------------------------------------------------------------------------------
# a13: halo exchange — src/parallelization_utils.py:6-52
# ---------------------------------------------------------------------------------------------------------
def self_exchange(f):
    """One rank: the four Sendrecv calls are self-copies in the order left, right, bottom, top (:34-49);
    the y faces go after the x faces and so carry the x-ghost corners."""
    f[-1, :, :] = f[1, :, :]
    f[0, :, :] = f[-2, :, :]
    f[:, -1, :] = f[:, 1, :]
    f[:, 0, :] = f[:, -2, :]
    return f


def exchange_blocks(blocks, xs, ys):
    """k ranks emulated in one process: blocks[(cx,cy)] are ghost-padded local arrays; same four phases,
    each completed on all ranks before the next (blocking Sendrecv semantics), periodic in both directions."""
    for (cx, cy), f in blocks.items():
        f[-1, :, :] = blocks[((cx + 1) % xs, cy)][1, :, :]       # everyone sends row 1 to the left
    for (cx, cy), f in blocks.items():
        f[0, :, :] = blocks[((cx - 1) % xs, cy)][-2, :, :]       # row -2 to the right
    for (cx, cy), f in blocks.items():
        f[:, -1, :] = blocks[(cx, (cy + 1) % ys)][:, 1, :]       # column 1 down
    for (cx, cy), f in blocks.items():
        f[:, 0, :] = blocks[(cx, (cy - 1) % ys)][:, -2, :]       # column -2 up
    return blocks


# ---------------------------------------------------------------------------------------------------------
# a1: the step — src/lattice_boltzmann_method.py:191-228
# ---------------------------------------------------------------------------------------------------------
def step(f, rho, u, omega, boundary=None, comm=None):
    assert f.shape[0:2] == rho.shape and f.shape[0:2] == u.shape[0:2]
    assert 0 < omega < 2
    f_pre = collide(f, rho, u, omega)
    if comm is not None:
        f_pre = comm(f_pre)
    f_post = stream(f_pre)
    if boundary is not None:
        f_post = boundary(f_pre, f_post, rho, u, f)
    rho2 = density(f_post)
    return f_post, rho2, velocity(rho2, f_post)


def step_blocks(F, R, U, omega, bcs, xs, ys):
    """One step of a k-rank run emulated in-process (dicts keyed by block coordinate)."""
    pre = {c: collide(F[c], R[c], U[c], omega) for c in F}
    exchange_blocks(pre, xs, ys)
    out_f, out_r, out_u = {}, {}, {}
    for c in F:
        fp = stream(pre[c])
        fp = bcs[c](pre[c], fp, R[c], U[c], F[c])
        out_f[c] = fp
        out_r[c] = density(fp)
        out_u[c] = velocity(out_r[c], fp)
    return out_f, out_r, out_u


# initial fields — src/initial_values.py:38-123 (host-side setup used by every driver)
def sinusoidal_velocity_x(shape, eps):
    """:67-93 — rho = 1, u_x(y) = eps*sin(2*pi*y/ly)."""
    rho = np.ones(shape)
    y = np.arange(shape[1])
    prof = eps * np.sin(np.divide(2 * np.pi * y, shape[1]))
    u = np.zeros(shape + (2,))
    u[..., 0] = prof[None, :]
    return rho, u


def sinusoidal_density_x(shape, p0, eps):
    """:38-64 — rho(x) = p0 + eps*sin(2*pi*x/lx), u = 0."""
    x = np.arange(shape[0])
    prof = p0 + eps * np.sin(np.divide(2 * np.pi * x, shape[0]))
    rho = np.repeat(prof[:, None], shape[1], axis=1)
    return rho, np.zeros(shape + (2,))


def uniform(shape, rho0=1.0, ux=0.0, uy=0.0):
    """:96-123."""
    u = np.empty(shape + (2,))
    u[..., 0] = ux
    u[..., 1] = uy
    return np.ones(shape) * rho0, u
