"""Host-side semantics of the lazy results of `lattice_boltzmann_step` (engine.py), on the CPU: the C library is
replaced by tests/fake_native.py (oracle arithmetic), everything above the C-ABI is the product code. What must
hold (SURVEY.md §8(b) "ownership"): the call is purely functional — results handed out earlier stay valid and
unchanged however far the resident lattice has advanced; handles fed back advance the lattice without transfers."""
import numpy as np
import pytest

from oracle import lbm_numpy as onp
from tests.fake_native import FakeLib


@pytest.fixture()
def L(monkeypatch):
    from lattice_boltzmann_parallel_solver_b200 import _native as N
    from lattice_boltzmann_parallel_solver_b200 import lattice_boltzmann_method as L
    fake = FakeLib()
    monkeypatch.setattr(N, 'load', lambda: fake)
    monkeypatch.setattr(N, 'device', lambda: 0)
    L._lattices.clear()
    L.fake = fake
    yield L
    L._lattices.clear()


def start(shape=(12, 10), seed=0):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(0.9, 1.1, shape)
    u = rng.uniform(-0.05, 0.05, shape + (2,))
    return onp.equilibrium(rho, u), rho, u


def test_loop_is_device_resident_and_matches_the_oracle(L):
    f, rho, u = start()
    ref = (f, rho, u)
    for _ in range(25):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.1)
        ref = onp.step(*ref, 1.1)
    lat = next(iter(L._lattices.values()))[0]
    assert lat.time == 0 and lat._pending_n == 25 and lat._pending == 1.1   # nothing launched until somebody looks
    assert len(L.fake.ctxs) == 1
    calls = list(L.fake._c(lat._ctx).calls)
    assert np.array_equal(np.asarray(u), ref[2])             # first access: ONE lbm_step(25) + one materialisation
    assert lat.time == 25 and L.fake._c(lat._ctx).calls == calls + [('step', 25), ('fields', 1)]
    assert np.array_equal(np.asarray(f), ref[0]) and np.array_equal(rho, ref[1])
    assert L.fake._c(lat._ctx).calls[-2:] == [('fields', 1), ('fields', 1)]   # u is cached, f and rho fetched once each


def test_kept_results_stay_valid(L):
    """experiments.py:254 appends every step's velocity to a list and reads them after the loop."""
    f, rho, u = start(seed=1)
    ref = (f, rho, u)
    kept, kept_ref = [], []
    for t in range(9):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 0.9)
        ref = onp.step(*ref, 0.9)
        if t % 2 == 0:
            kept.append(u)
            kept_ref.append(ref[2])
    for a, b in zip(kept, kept_ref):
        assert np.array_equal(np.asarray(a), b)
    assert np.array_equal(np.asarray(f), ref[0])


def test_inputs_are_never_modified_and_can_be_reused(L):
    """experiments.py:168-176 reuses one initial tuple for all omegas."""
    f0, rho0, u0 = start(seed=2)
    keep = (f0.copy(), rho0.copy(), u0.copy())
    outs = []
    for om in (0.5, 1.0, 1.7):
        f, rho, u = f0, rho0, u0
        for _ in range(4):
            f, rho, u = L.lattice_boltzmann_step(f, rho, u, om)
        outs.append((om, f, rho, u))
    for a, b in zip((f0, rho0, u0), keep):
        assert np.array_equal(a, b)
    for om, f, rho, u in outs:                               # all three runs' results are still readable
        ref = keep
        for _ in range(4):
            ref = onp.step(*ref, om)
        assert np.array_equal(np.asarray(f), ref[0]) and np.array_equal(np.asarray(u), ref[2])


def test_cell_reads_extrema_and_ndarray_behaviour(L):
    f, rho, u = start(seed=3)
    ref = (f, rho, u)
    for _ in range(3):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.2)
        ref = onp.step(*ref, 1.2)
        assert np.array_equal(np.array(u[4, 5, ...]), ref[2][4, 5])     # experiments.py:703
        assert np.linalg.norm(u[4, 5, ...]) == np.linalg.norm(ref[2][4, 5])
        assert np.amin(rho) == ref[1].min() and np.amax(u) == ref[2].max()   # experiments.py:181-193
        assert u._value is None and rho._value is None                   # none of that materialised a field
    assert u.shape == (12, 10, 2) and rho.ndim == 2 and len(f) == 12 and f.dtype == np.float64
    assert np.array_equal(u[..., 0], ref[2][..., 0])                     # experiments.py:326
    assert np.array_equal(u[1:-1, 1:-1, :], ref[2][1:-1, 1:-1, :])       # experiments.py:639
    assert np.allclose((rho * 2 + 1).sum(), (ref[1] * 2 + 1).sum())
    assert np.array_equal(np.stack([rho, rho]), np.stack([ref[1], ref[1]]))
    assert float(rho.mean()) == float(ref[1].mean()) and rho.copy().flags.writeable
    with pytest.raises(ValueError):
        np.asarray(rho)[0, 0] = 1.0                                      # cached results are read-only views


def test_per_step_cell_reads_move_to_the_probe_ring(L):
    """experiments.py:703-704 reads velocity[px, py] after EVERY step: the second such read configures the probe on
    that cell, later reads are served from the ring (no materialisation launch); other cells, other fields, a jump
    in time and a fresh upload fall back to the region read."""
    f, rho, u = start(seed=7)
    ref = (f, rho, u)
    lat = None
    for t in range(1, 9):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.3)
        ref = onp.step(*ref, 1.3)
        assert np.array_equal(np.array(u[4, 5, ...]), ref[2][4, 5])
        assert u[4, 5, 1] == ref[2][4, 5, 1] and rho[4, 5] == ref[1][4, 5]
        lat = next(iter(L._lattices.values()))[0]
        calls = L.fake._c(lat._ctx).calls
        if t == 2:
            assert ('probe_config', 4, 5) in calls
        if t >= 3:
            assert calls[-3][0] == 'probe_read' and calls[-2][0] == 'probe_read' and calls[-1] == ('fields', 1)   # u, u, rho
    assert np.array_equal(np.array(u[2, 3]), ref[2][2, 3])               # another cell: region read, the probe stays
    assert lat._probe[:2] == (4, 5)
    for _ in range(3):                                                   # three steps without a read: the ring has them all
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.3)
        ref = onp.step(*ref, 1.3)
    assert np.array_equal(np.array(u[4, 5]), ref[2][4, 5])
    f, rho, u = L.lattice_boltzmann_step(np.asarray(f), np.asarray(rho), np.asarray(u), 1.3)   # fresh upload
    ref = onp.step(*ref, 1.3)
    assert np.array_equal(np.array(u[4, 5]), ref[2][4, 5])


def test_a_probe_the_caller_configured_is_used_and_kept(L):
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    f, rho, u = start(seed=8)
    lat = Lattice(12, 10)
    lat.probe(3, 3, capacity=16)
    lat.load(f, rho, u, 1.0)
    ref = (f, rho, u)
    for t in range(1, 5):
        hs = lat.request_step(1.0)
        ref = onp.step(*ref, 1.0)
        assert np.array_equal(np.array(hs[2][3, 3]), ref[2][3, 3])       # on the probe cell: from the ring
        assert np.array_equal(np.array(hs[2][6, 2]), ref[2][6, 2])       # elsewhere: region reads, no re-configuration
    calls = L.fake._c(lat._ctx).calls
    assert [c for c in calls if c[0] == 'probe_config'] == [('probe_config', 3, 3)]
    assert sum(1 for c in calls if c[0] == 'probe_read') == 4


def test_modified_result_forces_a_fresh_upload(L):
    f, rho, u = start(seed=4)
    f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.0)
    ref = onp.step(*start(seed=4), 1.0)
    rho[0, 0] = 1.5                                                      # tracked edit through __setitem__
    ref[1][0, 0] = 1.5
    f2, rho2, u2 = L.lattice_boltzmann_step(f, rho, u, 1.0)
    exp = onp.step(ref[0], ref[1], ref[2], 1.0)
    assert np.array_equal(np.asarray(f2), exp[0])
    lat = next(iter(L._lattices.values()))[0]
    assert lat.time == 1                                                 # re-uploaded: the lattice restarted its clock


def test_mixed_and_foreign_arguments_upload(L):
    f, rho, u = start(seed=5)
    a = L.lattice_boltzmann_step(f, rho, u, 1.0)
    b = L.lattice_boltzmann_step(a[0], np.asarray(a[1]), a[2], 1.0)      # one plain array among handles
    ref = onp.step(*onp.step(f, rho, u, 1.0), 1.0)
    assert np.array_equal(np.asarray(b[0]), ref[0])
    old = L.lattice_boltzmann_step(f, rho, u, 1.0)                       # handles of a restarted lattice ...
    c = L.lattice_boltzmann_step(*b, 1.0)                                # ... and stale-but-materialised ones still work
    assert np.array_equal(np.asarray(c[2]), onp.step(*ref, 1.0)[2])
    assert np.array_equal(np.asarray(old[1]), onp.step(f, rho, u, 1.0)[1])


def test_omega_per_call(L):
    f, rho, u = start(seed=6)
    ref = (f, rho, u)
    for om in (0.3, 0.3, 1.9, 1.0):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, om)
        ref = onp.step(*ref, om)
    assert np.array_equal(np.asarray(f), ref[0])
    lat = next(iter(L._lattices.values()))[0]
    assert L.fake._c(lat._ctx).steps_log == [0.3, 0.3, 1.9, 1.0]


def test_direct_native_advance_makes_old_handles_loudly_stale(L):
    f, rho, u = start(seed=7)
    f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.0)
    lat = next(iter(L._lattices.values()))[0]
    lat.flush()
    lat.run(3)                                                           # bypasses the handle protocol
    with pytest.raises(RuntimeError, match='stale'):
        np.asarray(u)


def test_cache_eviction_and_release_keep_results_readable(L):
    results = []
    for n in range(11):                                  # more lattice shapes than the cache holds
        f, rho, u = start(shape=(6 + n, 5), seed=n)
        out = L.lattice_boltzmann_step(f, rho, u, 1.0)
        out = L.lattice_boltzmann_step(*out, 1.0)
        results.append((out, onp.step(*onp.step(f, rho, u, 1.0), 1.0)))
    assert len(L._lattices) <= 9
    L.release_lattices()
    assert not L._lattices and not L.fake.ctxs           # every device lattice is gone ...
    for out, ref in results:                             # ... and every result handed out is still right
        for a, b in zip(out, ref):
            assert np.array_equal(np.asarray(a), b)


def test_cell_index_bounds(L):
    f, rho, u = start(seed=9)
    f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.0)
    ref = onp.step(*start(seed=9), 1.0)
    assert np.array_equal(np.array(u[-1, -2]), ref[2][-1, -2])
    with pytest.raises(IndexError):
        u[12, 0]


def test_long_loops_are_launched_in_batches(L):
    """What lets the reference's driver loops reach the two-steps-per-pass kernel / graph replay: lbm_step(n) with
    n = MAX_DEFERRED, not n = 1."""
    from lattice_boltzmann_parallel_solver_b200.engine import MAX_DEFERRED
    f, rho, u = start(seed=11)
    ref = (f, rho, u)
    n = 3 * MAX_DEFERRED + 5
    for _ in range(n):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.3)
        ref = onp.step(*ref, 1.3)
    lat = next(iter(L._lattices.values()))[0]
    steps = [c for c in L.fake._c(lat._ctx).calls if c[0] == 'step']
    assert steps == [('step', MAX_DEFERRED)] * 3
    assert np.array_equal(np.asarray(rho), ref[1])
    assert [c for c in L.fake._c(lat._ctx).calls if c[0] == 'step'][-1] == ('step', 5)


def test_results_kept_in_the_middle_of_a_batch(L):
    f, rho, u = start(seed=12)
    ref = (f, rho, u)
    kept = {}
    for t in range(1, 41):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 0.7)
        ref = onp.step(*ref, 0.7)
        if t in (7, 8, 23):
            kept[t] = (rho, ref[1].copy())
        if t == 30:
            assert np.array_equal(np.array(u[3, 3]), ref[2][3, 3])       # a look in the middle of the loop
    lat = next(iter(L._lattices.values()))[0]
    # the batch was split exactly where somebody still held a result: 7, 8, 23, then the look at 30
    assert [c[1] for c in L.fake._c(lat._ctx).calls if c[0] == 'step'] == [7, 1, 15, 7]
    for t, (h, want) in kept.items():
        assert np.array_equal(np.asarray(h), want), t
    assert np.array_equal(np.asarray(f), ref[0]) and lat.time == 40


def test_old_handle_read_while_newer_steps_are_queued(L):
    f, rho, u = start(seed=13)
    ref = (f, rho, u)
    hist = []
    for t in range(10):
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.0)
        ref = onp.step(*ref, 1.0)
        hist.append((u, ref[2]))
    lat = next(iter(L._lattices.values()))[0]
    assert np.array_equal(np.asarray(hist[3][0]), hist[3][1])            # advances the device to time 4 only
    assert lat.time == 4 and lat._pending_n == 6
    assert np.array_equal(np.asarray(hist[9][0]), hist[9][1]) and lat.time == 10
    for h, want in hist:
        assert np.array_equal(np.asarray(h), want)


def test_results_can_be_edited_in_place_like_the_reference_arrays(L):
    """The reference hands back ordinary writable ndarrays: `f += x`, np.add(f, x, out=f) must work on a handle."""
    f, rho, u = start(seed=11)
    hf, hr, hu = L.lattice_boltzmann_step(f, rho, u, 1.0)
    ref = [a.copy() for a in onp.step(f, rho, u, 1.0)]
    keep = hf
    hf += 0.25
    assert hf is keep
    ref[0] += 0.25
    np.multiply(hr, 2.0, out=hr)
    ref[1] *= 2.0
    assert np.array_equal(np.asarray(hf), ref[0]) and np.array_equal(np.asarray(hr), ref[1])
    assert not np.asarray(hu).flags.writeable            # documented: the plain host copy is read-only
    nxt = L.lattice_boltzmann_step(hf, hr, hu, 1.0)      # edited handles are uploaded, not taken for the device state
    exp = onp.step(ref[0], ref[1], ref[2], 1.0)
    assert np.array_equal(np.asarray(nxt[0]), exp[0]) and np.array_equal(np.asarray(nxt[2]), exp[2])


def test_a_sweep_does_not_pin_device_lattices(L):
    """Every experiment of a sweep builds fresh closures (src/experiments.py:692-694, 759-761): identical scenarios
    must reuse one device lattice, and lattices nobody reads from any more must be freed before new ones are made."""
    for rep in range(3):                                  # same scenario three times: one context
        f, rho, u = start((12, 10), seed=rep)
        for _ in range(3):
            f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.0)
        assert np.array_equal(np.asarray(rho), _run_ref((12, 10), rep, 3)[1])
    assert len(L.fake.ctxs) == 1
    for n in range(13, 20):                               # a sweep over sizes: old lattices are retired
        f, rho, u = start((n, 10), seed=n)
        f, rho, u = L.lattice_boltzmann_step(f, rho, u, 1.0)
        np.asarray(u)
        del f, rho, u
    assert len(L.fake.ctxs) <= L.MAX_IDLE_LATTICES + 1
    # a lattice somebody still holds an unread result of is never retired behind their back
    f, rho, u = start((30, 10), seed=1)
    held = L.lattice_boltzmann_step(f, rho, u, 1.0)
    for n in range(31, 36):
        g = L.lattice_boltzmann_step(*start((n, 10), seed=n), 1.0)
        np.asarray(g[0])
        del g
    assert np.array_equal(np.asarray(held[2]), onp.step(f, rho, u, 1.0)[2])


def _run_ref(shape, seed, n):
    s = start(shape, seed)
    for _ in range(n):
        s = onp.step(*s, 1.0)
    return s


def test_out_of_device_memory_retires_idle_lattices_and_retries(L):
    real_create = L.fake.lbm_create

    def create(device, nx, ny, gx, gy, bc, out):
        if len(L.fake.ctxs) >= 2:
            L.fake.err = b'cannot allocate'
            return 4                                      # LBM_ERR_NOMEM
        return real_create(device, nx, ny, gx, gy, bc, out)
    L.fake.lbm_create = create
    for n in (12, 13, 14, 15):
        g = L.lattice_boltzmann_step(*start((n, 10), seed=n), 1.0)
        assert np.array_equal(np.asarray(g[1]), _run_ref((n, 10), n, 1)[1])
    assert len(L.fake.ctxs) <= 2


def test_run_host_is_load_run_fields_and_keeps_handles_valid(L):
    """Lattice.run_host (lbm_run_host: the whole job from and to host arrays in one call) must leave results handed out
    earlier readable and restart the lattice's clock like a load."""
    from lattice_boltzmann_parallel_solver_b200.engine import Lattice
    f, rho, u = start(seed=21)
    lat = Lattice(12, 10)
    lat.load(f, rho, u, 1.0)
    old = lat.request_step(1.0)                          # a queued step whose result somebody still holds
    g = start(seed=22)
    out = lat.run_host(*g, 1.3, 5, out=(g[0].copy(), None, g[2].copy()))
    ref = g
    for _ in range(5):
        ref = onp.step(*ref, 1.3)
    assert np.array_equal(out[0], ref[0]) and out[1] is None and np.array_equal(out[2], ref[2])
    assert lat.time == 5 and lat._pending_n == 0
    assert np.array_equal(np.asarray(old[1]), onp.step(f, rho, u, 1.0)[1])      # preserved before the lattice was reloaded
    nxt = lat.request_step(1.3)                          # ... and the lattice goes on from the state run_host left
    assert np.array_equal(np.asarray(nxt[2]), onp.step(*ref, 1.3)[2])


def test_kept_velocities_are_parked_on_the_device(L):
    """velocities.append(velocity) every step (experiments.py:254, :542): no host materialisation inside the loop; the
    fields come out of the device history when (and only when) they are looked at, and slots are recycled."""
    lib, P = L.fake, L
    f, rho, u = start(seed=5)
    rf, rr, ru = f, rho, u
    want = []
    kept, kept_rho = [], []
    for i in range(20):
        f, rho, u = P.lattice_boltzmann_step(f, rho, u, 1.1)
        kept.append(u)
        if i % 5 == 0:
            kept_rho.append(rho)
        rf, rr, ru = onp.step(rf, rr, ru, 1.1)
        want.append((rr.copy(), ru.copy()))
    c = next(iter(lib.ctxs.values()))
    np.testing.assert_array_equal(np.asarray(u), want[-1][1])           # flushes the queue
    assert ('fields', 1) not in c.calls[:-1], c.calls                    # nothing came to the host inside the loop
    assert sum(1 for x in c.calls if x[0] == 'hist_store') == 19
    assert 'parked' in repr(kept[3])
    np.testing.assert_array_equal(kept[3], want[3][1])
    np.testing.assert_array_equal(kept_rho[1], want[5][0])
    np.testing.assert_array_equal(np.asarray(kept[5]), want[5][1])       # same slot, other field
    assert sum(1 for x in c.calls if x[0] == 'hist_read') == 3
    lat = kept[0]._lattice
    n_free = len(lat._hist_free)
    del kept[11:15]      # (the slot of step 10 is still held by its density handle)
    import gc; gc.collect()
    assert len(lat._hist_free) == n_free + 4                             # dropped handles give their slots back
    # a new upload on the same lattice (other initial state) leaves parked results readable; closing drains them
    np.testing.assert_array_equal(kept[7], want[7][1])
    lat.retire()
    for i, j in ((0, 0), (1, 1), (2, 2), (4, 4), (6, 6), (8, 8), (9, 9), (10, 10), (11, 15), (12, 16), (14, 18)):
        np.testing.assert_array_equal(kept[i], want[j][1])


def test_history_falls_back_to_the_host_without_slots(L, monkeypatch):
    lib, P = L.fake, L
    from lattice_boltzmann_parallel_solver_b200 import engine
    monkeypatch.setattr(engine, 'HISTORY_MAX_SLOTS', 4)
    f, rho, u = start((8, 8), seed=6)
    rf, rr, ru = f, rho, u
    kept, want = [], []
    for i in range(9):
        f, rho, u = P.lattice_boltzmann_step(f, rho, u, 0.9)
        kept.append(u)
        rf, rr, ru = onp.step(rf, rr, ru, 0.9)
        want.append(ru.copy())
    np.testing.assert_array_equal(kept[-1], want[-1])       # the device passes steps 1..8, all of them still referenced
    c = next(iter(lib.ctxs.values()))
    assert sum(1 for x in c.calls if x[0] == 'hist_store') == 4 and c.calls.count(('fields', 1)) == 5
    for a, b in zip(kept, want):
        np.testing.assert_array_equal(a, b)


def test_parked_handles_behave_like_arrays(L):
    """A result parked in the device history is still an ndarray to its owner: cell reads, whole-field reductions, in-place
    edits (private copy, re-upload when fed back) all work on it after the device has moved on."""
    lib, P = L.fake, L
    f, rho, u = start(seed=9)
    rf, rr, ru = f, rho, u
    kept_u, kept_rho, want = [], [], []
    for i in range(6):
        f, rho, u = P.lattice_boltzmann_step(f, rho, u, 1.2)
        kept_u.append(u)
        kept_rho.append(rho)
        rf, rr, ru = onp.step(rf, rr, ru, 1.2)
        want.append((rf.copy(), rr.copy(), ru.copy()))
    np.asarray(f)                                               # the device is at step 6; steps 1-5 are parked
    assert all(h._hist is not None for h in kept_u[:5] + kept_rho[:5])
    assert np.array_equal(kept_u[1][3, 4, ...], want[1][2][3, 4]) and kept_rho[2][5, 6] == want[2][1][5, 6]
    assert np.amax(kept_rho[3]) == want[3][1].max() and kept_u[3].min() == want[3][2].min()
    assert (kept_u[0] + 1.0).shape == (12, 10, 2) and np.array_equal(kept_u[0] * 2.0, want[0][2] * 2.0)
    h = kept_u[4]
    h[0, 0, 0] = 7.0                                            # in-place edit of a parked result: a private host copy
    assert h._dirty and h[0, 0, 0] == 7.0 and np.array_equal(np.asarray(h)[1:], want[4][2][1:])
    # feeding an old, edited state back uploads it (the reference's arrays are plain values)
    edited_u = np.asarray(h).copy()
    f2, rho2, u2 = P.lattice_boltzmann_step(np.asarray(want[4][0]), kept_rho[4], h, 1.2)
    ref2 = onp.step(want[4][0], want[4][1], edited_u, 1.2)
    assert np.array_equal(np.asarray(f2), ref2[0]) and np.array_equal(np.asarray(u2), ref2[2])
