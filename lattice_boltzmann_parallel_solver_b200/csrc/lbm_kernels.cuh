// lbm_kernels.cuh — the device side of liblbm_b200.so: kernel parameter blocks, the per-cell pieces shared by all
// kernels (pulls, boundary rules, ghost stores, cross-GPU flags), and the kernels themselves:
//   k_step<MASK,HALO,FINAL,LIST>   one step, one cell per thread (boundary rules, ghost stores, materialisation, edge list)
//   k_step_pair                    one step, two cells per thread: the 144 B-per-update bandwidth kernel
//   k_step2x<T,HALO,PROBE>         two steps per pass (round 1)
//   k_stepNx<T,D,HALO,PROBE,FINAL> D = 2..4 steps per pass: the headline kernel is D = 3
//   k_cluster_steps<MASK,M,TMAX>   whole lattice in one thread-block cluster's shared memory, many steps per launch
//   k_first_collide, stateless operators, min/max, the arithmetic self-test
// Included by lbm_b200.cu only (host side: contexts, scheduling, C-ABI). Arithmetic: lbm_device.cuh.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <cstdint>
#include <cstring>

#include "../../include/lbm_b200.h"
#include "lbm_device.cuh"

using namespace lbm;

// -------------------------------------------------------------------------------------------------------
// kernel parameter blocks
// -------------------------------------------------------------------------------------------------------
struct HaloTarget {
    double *base;      // population-0 plane of the neighbour's DESTINATION buffer (this step's parity); null = none
    long long plane;   // its plane stride (doubles)
    int nx, ny, pitch;
};

struct StepParams {
    const double *src;
    double *dst;
    long long plane;   // plane stride of src: NX * pitch
    long long dplane;  // plane stride of dst (differs from `plane` only when one side is a strip window)
    int sbase, dbase;  // strip windows (two_steps on BC lattices): buffer row of lattice row x is (x - base) mod NX
    int pitch, NX, NY;
    int gx, gy;
    // rows handled by this launch: [row0a, row0a+na) then [row0b, ...)
    int row0a, na, row0b;
    int bpr;           // blocks per row
    int seg, nb;       // k_step2x: output rows per block; length of the second row range
    int pf;            // k_step2x: rows ahead of the march whose source segments are prefetched into L2 (0 = off)
    int strip0;        // k_stepNx: index of the first column strip of this launch (materialising a sub-rectangle)
    int y0, y1;        // columns handled: [y0, y1)
    double omega;
    double omega_last; // k_stepNx: omega of the LAST level's collision (differs when the caller changed omega)
    const uint8_t *kind_map;   // [x*pitch + y] or null
    const lbm_kind *kinds;
    const double *ktab, *ctab;
    const double *out_cur;     // outlet side buffer read by OUTLET rules  [3][pitch]
    double *out_next;          // written by OUTLET_SRC cells
    double rho_in, rho_out;
    int px, py;
    double *probe;             // ring of (ux, uy), probe_cap entries, in host-mapped memory, or null
    long long *progress;       // host-mapped: time of the newest sample in the ring (lbm_probe_read polls it)
    const long long *tc_in;    // device-resident time of the state read (captured graphs need no new params) ...
    long long *tc_out;         // ... and of the state written; both point into lbm_ctx::tcount
    int probe_cap;
    // FINAL (materialize) outputs, packed over [ox0,ox1) x [oy0,oy1)
    double *o_f, *o_rho, *o_u;
    int ox0, oy0, ow;          // ow = oy1 - oy0
    // halo
    HaloTarget halo[9];
    // ghost snapshot (see snapshot_ghosts): [2][9][pitch] for rows 0 / NX-1, [2][9][NX] for columns 0 / NY-1
    double *snap_row, *snap_col;
    int use_snap;              // FINAL launches: read ghost cells from the snapshot instead of S
    int no_snap;               // launches through a strip window: take no ghost snapshot (the source is not S)
    int eager_progress;        // one-step call (a driver that reads the probe cell after EVERY step): publish the time word in this launch
    // fix-up list
    const int2 *cells;
    int n_cells;
    // cross-GPU step flags
    volatile unsigned *flag_in;     // [9] in my arena: neighbour slot s finished writing my ghosts of step value
    unsigned *flag_out[9];          // neighbour's flag_in[opposite slot] (peer memory) or null
    unsigned wait_value, signal_value;
    long long timeout_cycles;
    unsigned *done_counter;         // last-block-done counter
    unsigned *err_flag;
    unsigned *err_host;             // host-mapped mirror of err_flag
    int n_blocks;
};

// -------------------------------------------------------------------------------------------------------
// per-cell pieces shared by all kernels
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ldS(const double *p) { return __ldcg(p); }   // L2-coherent (peer-written ghosts)

// Buffer row of lattice row x. Whole-lattice buffers have base 0; a strip window (see two_steps) holds the rows
// base, base+1, ... (mod NX) of the lattice in its rows 0, 1, ...
__device__ __forceinline__ long long srow(const StepParams &P, int x) { return x - P.sbase + (x < P.sbase ? P.NX : 0); }
__device__ __forceinline__ long long drow(const StepParams &P, int x) { return x - P.dbase + (x < P.dbase ? P.NX : 0); }

// Source value S[i][xs][ys]; materialisation launches take ghost cells from the snapshot (see snapshot_ghosts).
__device__ __forceinline__ double ld_cell(const StepParams &P, int i, int xs, int ys)
{
    if (P.use_snap) {
        // the ghost row next to the interior (row gx-1 / NX-gx) is the only one an interior cell ever pulls from
        if (P.gx && (xs == P.gx - 1 || xs == P.NX - P.gx)) return P.snap_row[((xs >= P.gx ? 1 : 0) * 9 + i) * (long long)P.pitch + ys];
        if (P.gy && (ys == 0 || ys == P.NY - 1)) return P.snap_col[((ys ? 1 : 0) * 9 + i) * (long long)P.NX + xs];
    }
    return ldS(P.src + i * P.plane + srow(P, xs) * P.pitch + ys);
}

// Values of time t are reconstructed from S_{t-1}, ghost cells included. A neighbour that is one step ahead
// overwrites MY ghost cells of that buffer with its S_{t+1} as soon as I have finished step t — possibly before I
// materialise time t. So the step kernel keeps a private copy of the ghost ring of the buffer it reads (edge
// threads copy the ghost cells next to them; ~2(NX+NY) cells), and materialisation reads ghosts from that copy.
__device__ __forceinline__ void snapshot_ghosts(const StepParams &P, int x, int y)
{
    const bool xl = P.gx && x == P.gx, xh = P.gx && x == P.NX - 1 - P.gx;
    const bool yl = P.gy && y == P.gy, yh = P.gy && y == P.NY - 1 - P.gy;
    if (!(xl | xh | yl | yh)) return;
    // All loads of a copy are issued before its first store: the compiler cannot prove that the snapshot does not
    // alias S, and a load-store-load-store chain costs nine L2 round trips per copy (this was most of a
    // launch-bound von Karman step, profiles/r01_summary.md section 9).
    auto copy9 = [&](const double *src, long long sstride, double *dst, long long dstride) {
        double v[9];
#pragma unroll
        for (int i = 0; i < 9; i++) v[i] = ldS(src + i * sstride);
#pragma unroll
        for (int i = 0; i < 9; i++) dst[i * dstride] = v[i];
    };
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        if (side ? xh : xl) {
            const int gx_row = side ? P.NX - P.gx : P.gx - 1;
            const double *src = P.src + (long long)gx_row * P.pitch;
            double *dst = P.snap_row + (long long)side * 9 * P.pitch;
            copy9(src + y, P.plane, dst + y, P.pitch);
            if (yl) copy9(src, P.plane, dst, P.pitch);
            if (yh) copy9(src + P.NY - 1, P.plane, dst + P.NY - 1, P.pitch);
        }
        if (side ? yh : yl) {
            const int gy_col = side ? P.NY - 1 : 0;
            copy9(P.src + (long long)x * P.pitch + gy_col, P.plane, P.snap_col + (long long)side * 9 * P.NX + x, P.NX);
        }
    }
}

// f_post[I] of a non-fluid cell by its rule (rule table of include/lbm_b200.h). I is a compile-time index so that
// the nine results stay in registers: no out-of-line call, no local-memory array anywhere in the step kernels.
template <int I>
__device__ __forceinline__ double pull_rule(const StepParams &P, const lbm_kind &k, int x, int y)
{
    constexpr int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    constexpr int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    const long long pl = P.plane;
    const int r = k.rule[I], type = r & 7, row = r >> 3;
    if (type == LBM_RULE_PULL) {
        int xs = x - cx[I], ys = y - cy[I];
        xs = xs < 0 ? P.NX - 1 : (xs >= P.NX ? 0 : xs);
        ys = ys < 0 ? P.NY - 1 : (ys >= P.NY ? 0 : ys);
        return ld_cell(P, I, xs, ys);
    }
    if (type == LBM_RULE_BOUNCE) {
        const double v = ldS(P.src + opp[I] * pl + srow(P, x) * P.pitch + y);
        return row ? sub(v, P.ktab[row * 9 + opp[I]]) : v;
    }
    if (type == LBM_RULE_CONST) return P.ctab[row * 9 + I];
    // LBM_RULE_OUTLET: populations 3, 6, 7 -> slots 0, 1, 2 of the side buffer
    return ldS(P.out_cur + (I == 3 ? 0 : (I == 6 ? 1 : 2)) * P.pitch + y);   // (L2: written one step earlier, possibly by the same launch)
}

__device__ __forceinline__ void pull_rules(const StepParams &P, const lbm_kind &k, int x, int y, double (&f)[9])
{
    f[0] = pull_rule<0>(P, k, x, y);
    f[1] = pull_rule<1>(P, k, x, y);
    f[2] = pull_rule<2>(P, k, x, y);
    f[3] = pull_rule<3>(P, k, x, y);
    f[4] = pull_rule<4>(P, k, x, y);
    f[5] = pull_rule<5>(P, k, x, y);
    f[6] = pull_rule<6>(P, k, x, y);
    f[7] = pull_rule<7>(P, k, x, y);
    f[8] = pull_rule<8>(P, k, x, y);
}

// Everything after the collision: stores of S' (own cell, PBC-owned virtual cells, neighbours' ghosts)
__device__ __forceinline__ void store_cell(const StepParams &P, int x, int y, const double (&s)[9], unsigned skip)
{
    double *d = P.dst + drow(P, x) * P.pitch + y;
    const long long pl = P.dplane;
    if (skip == 0) {
#pragma unroll
        for (int i = 0; i < 9; i++) __stcg(d + i * pl, s[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++)
            if (!((skip >> i) & 1)) __stcg(d + i * pl, s[i]);
    }
}

// periodic_with_pressure_variations (boundary_conditions.py:337-344): the cell on row -2 (resp. 1) produces the
// pre-streaming populations of the virtual node on row 0 (resp. -1) for the NEXT step:
//   feq_d(rho_b, u) + (f_pre_d - feq_d(rho, u)),  f_pre_d = the value just collided (s), u/rho this cell's moments
__device__ __forceinline__ void store_pbc(const StepParams &P, unsigned flags, int y, const double (&s)[9],
                                       const double (&p)[9], const double (&e)[9])
{
    const long long pl = P.dplane;   // (never launched on strip windows: lattices with this boundary have no clean rows)
    if (flags & LBM_CELL_PBC_IN_SRC) {
        const double w1 = mul(LBM_W1, P.rho_in), w5 = mul(LBM_W5, P.rho_in);
        double *d = P.dst + y;  // row 0
        __stcg(d + 1 * pl, add(mul(w1, p[1]), sub(s[1], e[1])));
        __stcg(d + 5 * pl, add(mul(w5, p[5]), sub(s[5], e[5])));
        __stcg(d + 8 * pl, add(mul(w5, p[8]), sub(s[8], e[8])));
    }
    if (flags & LBM_CELL_PBC_OUT_SRC) {
        const double w1 = mul(LBM_W1, P.rho_out), w5 = mul(LBM_W5, P.rho_out);
        double *d = P.dst + (long long)(P.NX - 1) * P.pitch + y;  // row -1
        __stcg(d + 3 * pl, add(mul(w1, p[3]), sub(s[3], e[3])));
        __stcg(d + 6 * pl, add(mul(w5, p[6]), sub(s[6], e[6])));
        __stcg(d + 7 * pl, add(mul(w5, p[7]), sub(s[7], e[7])));
    }
}

// communication() (parallelization_utils.py:34-49) without the copy: an interior cell on the edge of the block
// also writes its nine post-collision populations into the ghost cell(s) of the neighbour(s) that mirror it.
__device__ __forceinline__ void store_halo(const StepParams &P, int x, int y, const double (&s)[9])
{
    // an interior cell within gx rows of a block edge mirrors into the neighbour's ghost row at the same depth
    const int ex_lo = (P.gx && x < 2 * P.gx), ex_hi = (P.gx && x >= P.NX - 2 * P.gx);
    const int ey_lo = (P.gy && y == P.gy), ey_hi = (P.gy && y == P.NY - 1 - P.gy);
    if (!(ex_lo | ex_hi | ey_lo | ey_hi)) return;
#pragma unroll 1
    for (int ix = 0; ix < 3; ix++) {          // ix: 0 -> neighbour at dx=-1, 1 -> same, 2 -> dx=+1
        if ((ix == 0 && !ex_lo) || (ix == 2 && !ex_hi)) continue;
#pragma unroll 1
        for (int iy = 0; iy < 3; iy++) {
            if ((iy == 0 && !ey_lo) || (iy == 2 && !ey_hi)) continue;
            if (ix == 1 && iy == 1) continue;
            const HaloTarget &T = P.halo[ix * 3 + iy];
            if (!T.base) continue;
            // my first interior row is the low neighbour's high ghost row, and so on
            const int tx = ix == 0 ? T.nx - 2 * P.gx + x : (ix == 2 ? x - (P.NX - 2 * P.gx) : x);
            const int ty = iy == 0 ? T.ny - 1 : (iy == 2 ? 0 : y);
            double *d = T.base + (long long)tx * T.pitch + ty;
#pragma unroll
            for (int i = 0; i < 9; i++) d[i * T.plane] = s[i];
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// cross-GPU ordering: wait until every remote neighbour has published `wait_value`, publish `signal_value`
// once the whole grid has finished its (peer) stores.
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool halo_wait(const StepParams &P)
{
    if (P.wait_value == 0) return true;
    __shared__ int halo_ok;
    if (threadIdx.x == 0) {
        // Sticky: once a wait has timed out on this context every later ghost-touching kernel gives up at once — it
        // neither computes from stale ghosts nor stores or publishes anything, so the neighbours time out as well
        // instead of consuming wrong values, and the host reports LBM_ERR_TIMEOUT from its next call.
        int ok = *(volatile unsigned *)P.err_flag == 0;
        const long long t0 = clock64();
        for (int s = 0; s < 9 && ok; s++) {
            if (!P.flag_out[s]) continue;   // not a remote neighbour
            while ((int)(P.flag_in[s] - P.wait_value) < 0) {
                if (clock64() - t0 > P.timeout_cycles) {   // a peer is not stepping in lockstep
                    const unsigned code = 0x80000000u | (P.wait_value << 8) | (unsigned)s;   // who waited for what
                    atomicCAS(P.err_flag, 0u, code);
                    *(volatile unsigned *)P.err_host = code;   // host-mapped mirror: the host sees it without a CUDA call
                    ok = 0;
                    break;
                }
                __nanosleep(200);
            }
        }
        __threadfence_system();
        halo_ok = ok;
    }
    __syncthreads();
    return halo_ok != 0;
}

__device__ __forceinline__ void halo_signal(const StepParams &P)
{
    if (P.signal_value == 0) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(P.done_counter, 1u);
        if (done == (unsigned)P.n_blocks - 1) {
            *P.done_counter = 0;
            __threadfence_system();
            for (int s = 0; s < 9; s++)
                if (P.flag_out[s]) *(volatile unsigned *)P.flag_out[s] = P.signal_value;
            __threadfence_system();
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// the fused step kernel
//   MASK : read the per-cell kind byte (boundary rules folded into the kernel)
//   HALO : edge cells also store into neighbours' ghost cells; flags ordering
//   FINAL: stop after the moments and write reference-layout f_post / rho / u (materialize)
//   LIST : cells come from a compact list (edge fix-up kernel) instead of a row range
// -------------------------------------------------------------------------------------------------------
// Probe (experiments.py:703-704): the one thread that owns the probe cell appends (ux, uy) of the new time to the
// ring and advances the device-side time counter of the destination buffer.
// The ring lives in host-mapped memory: the sample travels to the host as the step that produced it completes, and
// the host reads it without a CUDA call while the device runs on (lbm_probe_read). The sample(s) first, then a
// system-scope fence, then the time word the host polls.
__device__ __forceinline__ void publish_progress(const StepParams &P, long long t_new)
{
    __threadfence_system();
    *(volatile long long *)P.progress = t_new;
}

// One-step launches keep the system-scope fence off the step's critical path (it cost 0.9 us of a 4 us von Karman step):
// the sample of time t and the device clock are plain stores, and the time word the host polls is advanced by the
// probe thread of the NEXT launch — probe_prologue, right after the dependent-launch wait, i.e. when the launch that
// wrote the sample has completed and flushed. The newest sample of a drained stream needs no word at all
// (lbm_probe_read falls back to "both streams idle").
__device__ __forceinline__ long long probe_prologue(const StepParams &P)
{
    const long long t_in = __ldcg(P.tc_in);
    *(volatile long long *)P.progress = t_in;
    return t_in;
}

__device__ __forceinline__ void record_probe(const StepParams &P, double ux, double uy, long long t_in)
{
    const long long t_new = t_in + 1;
    double *slot = P.probe + 2 * (t_new % P.probe_cap);
    slot[0] = ux;
    slot[1] = uy;
    *P.tc_out = t_new;
    if (P.eager_progress) publish_progress(P, t_new);   // the host is waiting for exactly this sample: do not make it wait for the stream to drain
}

// Everything one cell does after its nine f_post values are known (shared by the register-resident fluid path
// and the out-of-line rule path).
template <bool HALO, bool FINAL>
__device__ __forceinline__ void finish_cell(const StepParams &P, int x, int y, const double (&f)[9], unsigned flags,
                                            unsigned skip, bool is_probe, long long t_in)
{
    // branch-free division / square root (lbm_device.cuh): on launch-bound lattices a step IS the dependent chain of one
    // cell update; operands the fast paths reject (never in a physical run) take the library operations
    double rho, ux, uy, p[9];
    bool slow = false;
    moments_fast(f, rho, ux, uy, slow);
    if (!FINAL) eq_poly_fast(ux, uy, p, slow);
    if (slow) {
        moments(f, rho, ux, uy);
        if (!FINAL) eq_poly(ux, uy, p);
    }
    if (FINAL) {
        const long long o = (long long)(x - P.ox0) * P.ow + (y - P.oy0);
        if (P.o_f) {
#pragma unroll
            for (int i = 0; i < 9; i++) P.o_f[o * 9 + i] = f[i];
        }
        if (P.o_rho) P.o_rho[o] = rho;
        if (P.o_u) {
            P.o_u[o * 2] = ux;
            P.o_u[o * 2 + 1] = uy;
        }
        return;
    }
    if (is_probe) record_probe(P, ux, uy, t_in);
    double e[9], s[9];
    eq_from_poly(rho, p, e);
    collide(f, e, P.omega, s);
    if (flags & LBM_CELL_OUTLET_SRC) {
        P.out_next[0 * P.pitch + y] = f[3];
        P.out_next[1 * P.pitch + y] = f[6];
        P.out_next[2 * P.pitch + y] = f[7];
    }
    store_cell(P, x, y, s, skip);
    if (flags & (LBM_CELL_PBC_IN_SRC | LBM_CELL_PBC_OUT_SRC)) store_pbc(P, flags, y, s, p, e);
    if (HALO) {
        store_halo(P, x, y, s);
        if (!P.no_snap) snapshot_ghosts(P, x, y);
    }
}

// Programmatic dependent launch (launch-bound lattices, where a step is ~1 us of work behind ~1 us of launch
// latency): a step kernel releases its dependents at once, so the next step's blocks are launched, read their
// parameters and their kind byte while this step still runs, and then wait here for this grid to complete and
// flush before they touch S. Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One cell of one step (or of a materialisation). Only the nine PULLS differ between a fluid cell and a non-fluid
// one (rule table, compile-time population indices: everything stays in registers); moments, collision and stores are
// one common instruction stream. (The first version sent non-fluid cells through their own copy of the whole cell
// update: a warp that holds a wall cell — two of the four warps of every Couette row — then ran the update twice, and
// config 2 cost 4.35 us per step against 2.0 us for the periodic lattice of config 1, profiles/r01_summary.md section 12.)
template <bool HALO, bool FINAL>
__device__ __forceinline__ void step_cell(const StepParams &P, int x, int y, unsigned kind, const lbm_kind &k)
{
    double f[9];
    unsigned flags = 0, skip = 0;
    const bool is_probe = !FINAL && P.probe && x == P.px && y == P.py;
    const long long t_in = is_probe ? probe_prologue(P) : 0;
    if (!(FINAL && P.use_snap)) {
        // ONE load per population for every kind of cell: the address is the streamed neighbour's (fluid cells, PULL
        // rules), the cell's own opposite population (bounce-back), the inlet constant in the table or the outlet side
        // buffer. A warp that holds boundary cells decodes their rules (integer work on a few lanes) and then issues
        // the same nine loads as every other warp — one L2 round trip instead of two (the rule lanes' pulls and the
        // fluid lanes' pulls used to run one after the other), which is what a step of a launch-bound lattice costs.
        constexpr int opp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
        const int xm = x == 0 ? P.NX - 1 : x - 1, xp = x == P.NX - 1 ? 0 : x + 1;
        const int ym = y == 0 ? P.NY - 1 : y - 1, yp = y == P.NY - 1 ? 0 : y + 1;
        const long long pl = P.plane;
        const double *r0 = P.src + srow(P, x) * P.pitch, *rm = P.src + srow(P, xm) * P.pitch, *rp = P.src + srow(P, xp) * P.pitch;
        const double *a[9] = {r0 + y,           rm + pl + y,      r0 + 2 * pl + ym, rp + 3 * pl + y, r0 + 4 * pl + yp,
                              rm + 5 * pl + ym, rp + 6 * pl + ym, rp + 7 * pl + yp, rm + 8 * pl + yp};
        double kv[9] = {};   // wall-velocity terms of moving_wall (boundary_conditions.py:207-210), 0 elsewhere
        if (kind != 0) {
            flags = k.flags;
            skip = k.skip_store;
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const int r = k.rule[i], type = r & 7, row = r >> 3;
                if (type == LBM_RULE_BOUNCE) {
                    a[i] = r0 + opp[i] * pl + y;
                    if (row) kv[i] = P.ktab[row * 9 + opp[i]];
                } else if (type == LBM_RULE_CONST) {
                    a[i] = P.ctab + row * 9 + i;
                } else if (type == LBM_RULE_OUTLET) {   // populations 3, 6, 7 -> slots 0, 1, 2 of the side buffer
                    a[i] = P.out_cur + (i == 3 ? 0 : (i == 6 ? 1 : 2)) * P.pitch + y;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 9; i++) f[i] = ldS(a[i]);
        if (kind != 0) {
#pragma unroll
            for (int i = 1; i < 9; i++) f[i] = sub(f[i], kv[i]);   // v - (+0.0) == v for every v
        }
    } else if (kind != 0) {   // materialisation of a block with ghost cells: ghosts come from the snapshot (ld_cell)
        pull_rules(P, k, x, y, f);
        flags = k.flags;
        skip = k.skip_store;
    } else {
        constexpr int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, cy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
#pragma unroll
        for (int i = 0; i < 9; i++) {
            int xs = x - cx[i], ys = y - cy[i];
            xs = xs < 0 ? P.NX - 1 : (xs >= P.NX ? 0 : xs);
            ys = ys < 0 ? P.NY - 1 : (ys >= P.NY ? 0 : ys);
            f[i] = ld_cell(P, i, xs, ys);
        }
    }
    finish_cell<HALO, FINAL>(P, x, y, f, flags, skip, is_probe, t_in);
}

template <bool MASK, bool HALO, bool FINAL, bool LIST>
__global__ void __launch_bounds__(256) k_step(const __grid_constant__ StepParams P)
{
    pdl_release();
    int x, y;
    bool active = true;
    if (LIST) {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        active = c < P.n_cells;
        const int2 xy = active ? P.cells[c] : make_int2(0, 0);
        x = xy.x;
        y = xy.y;
    } else {
        const int rb = blockIdx.x / P.bpr, cb = blockIdx.x - rb * P.bpr;
        x = rb < P.na ? P.row0a + rb : P.row0b + (rb - P.na);
        y = P.y0 + cb * blockDim.x + threadIdx.x;
        active = y < P.y1;
    }
    // the kind map and the kind table are written once, at lbm_create: safe to read before the previous step is complete
    unsigned kind = 0;
    lbm_kind k = {};
    if ((MASK || LIST) && active) {
        kind = P.kind_map[(long long)x * P.pitch + y];
        if (kind) k = P.kinds[kind];
    }
    pdl_wait();
    if (HALO && !FINAL && !halo_wait(P)) return;
    if (active) step_cell<HALO, FINAL>(P, x, y, kind, k);
    if (HALO && !FINAL) halo_signal(P);
}

// ==== HOT KERNELS BEGIN (bench.py hashes this region + lbm_device.cuh: profiles/traffic.json is quoted only for the
// ==== kernels it was captured on) ====
// -------------------------------------------------------------------------------------------------------
// The bandwidth kernel: fluid cells only (no kind byte, no ghost stores), TWO cells per thread along the fast axis.
//  * blockIdx.y is the row (no integer division), the three row bases are computed once per thread;
//  * the three populations that do not move along y (0, 1, 3) are read with 128-bit loads, the six shifted ones
//    with two 64-bit loads off one address register (+-1 element: 8 B alignment only);
//  * all nine populations of both cells are written with 128-bit stores.
// Same per-cell arithmetic as k_step (lbm_device.cuh), hence the same bits.
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ld2(const double *p)
{
    double2 v;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st2(double *p, double a, double b)
{
    asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}

// ILP: both cells of the pair in one basic block (branch-free arithmetic; 119 registers) — lattices whose step is bound by
// the latency of one cell update (up to 2^19 cells: 512^2 5.46 -> 5.31 us, 256^2 3.33 -> 3.01). Larger lattices keep the
// lean version (more resident warps: 1000^2 21.4 us vs 23.3 with ILP; 16384^2 5.66 ms either way).
template <bool ILP>
__global__ void __launch_bounds__(256) k_step_pair(const __grid_constant__ StepParams P)
{
    pdl_release();
    pdl_wait();
    const int x = P.row0a + blockIdx.y;
    const int y = 2 * (blockIdx.x * blockDim.x + threadIdx.x);   // cells y, y+1; NY is even
    if (y >= P.NY) return;
    const int xm = x == 0 ? P.NX - 1 : x - 1, xp = x == P.NX - 1 ? 0 : x + 1;
    const int ym = y == 0 ? P.NY - 1 : y - 1;          // left neighbour of the first cell
    const int yq = y + 2 == P.NY ? 0 : y + 2;          // right neighbour of the second cell
    const long long pl = P.plane;
    const double *r0 = P.src + (long long)x * P.pitch;
    const double *rm = P.src + (long long)xm * P.pitch;
    const double *rp = P.src + (long long)xp * P.pitch;
    const bool is_probe = P.probe && x == P.px && (y == P.py || y + 1 == P.py);
    const long long t_in = is_probe ? probe_prologue(P) : 0;

    double fa[9], fb[9];
    {
        const double2 v0 = ld2(r0 + y), v1 = ld2(rm + pl + y), v3 = ld2(rp + 3 * pl + y);
        fa[0] = v0.x; fb[0] = v0.y;
        fa[1] = v1.x; fb[1] = v1.y;
        fa[3] = v3.x; fb[3] = v3.y;
        // c_y = +1: cell y pulls from y-1, cell y+1 pulls from y
        fa[2] = ldS(r0 + 2 * pl + ym); fb[2] = ldS(r0 + 2 * pl + y);
        fa[5] = ldS(rm + 5 * pl + ym); fb[5] = ldS(rm + 5 * pl + y);
        fa[6] = ldS(rp + 6 * pl + ym); fb[6] = ldS(rp + 6 * pl + y);
        // c_y = -1: cell y pulls from y+1, cell y+1 pulls from y+2
        fa[4] = ldS(r0 + 4 * pl + y + 1); fb[4] = ldS(r0 + 4 * pl + yq);
        fa[7] = ldS(rp + 7 * pl + y + 1); fb[7] = ldS(rp + 7 * pl + yq);
        fa[8] = ldS(rm + 8 * pl + y + 1); fb[8] = ldS(rm + 8 * pl + yq);
    }
    // both cells in one basic block (branch-free division / square root, lbm_device.cuh): their dependent chains
    // interleave, which is most of what a step of a launch-bound lattice waits for
    double sa[9], sb[9], uax, uay, ubx, uby;
    if (ILP) {
        bool slow_a = false, slow_b = false;
        relax_fast(fa, P.omega, sa, uax, uay, slow_a);
        relax_fast(fb, P.omega, sb, ubx, uby, slow_b);
        if (slow_a | slow_b) {   // operands outside the fast paths' range: never in a physical run
            if (slow_a) relax_redo(fa, P.omega, sa, uax, uay);
            if (slow_b) relax_redo(fb, P.omega, sb, ubx, uby);
        }
    } else {
        {
            double rho, p[9], e[9];
            moments(fa, rho, uax, uay);
            eq_poly(uax, uay, p);
            eq_from_poly(rho, p, e);
            collide(fa, e, P.omega, sa);
        }
        {
            double rho, p[9], e[9];
            moments(fb, rho, ubx, uby);
            eq_poly(ubx, uby, p);
            eq_from_poly(rho, p, e);
            collide(fb, e, P.omega, sb);
        }
    }
    if (is_probe) record_probe(P, y == P.py ? uax : ubx, y == P.py ? uay : uby, t_in);
    double *d = P.dst + (long long)x * P.pitch + y;
#pragma unroll
    for (int i = 0; i < 9; i++) st2(d + i * pl, sa[i], sb[i]);
}

// -------------------------------------------------------------------------------------------------------
// TWO time steps per pass (temporal blocking) on fluid rows: S_t -> S_{t+2} with ~73 B of DRAM traffic per cell
// update instead of 144 B. A block of T threads owns output rows [x0, x1) x columns [y0, y0 + 2T-4) and marches
// along x; every thread owns an aligned PAIR of columns of the intermediate state S_{t+1}:
//   iteration j: pull row j of S_t from global (exactly the loads of k_step_pair; issued half an iteration ahead
//                into the registers row j-1 has just freed, after one thread per block has bulk-prefetched the
//                source segments of row j+2 into L2), collide -> row j of S_{t+1}. The six populations that move along y go into a 4-slot shared-memory
//                ring as 128-bit stores; the three that do not (0, 1, 3) never leave the thread's registers.
//                One __syncthreads; then row j-1 of S_{t+2} is pulled from ring rows j-2, j-1, j (the +-1 column
//                shifts are shared-memory offsets), collided and written with 128-bit stores.
// Redundant work: one intermediate pair each side of the strip (4/2T) and one intermediate row each end of the
// segment (2/seg). Same per-cell arithmetic as every other kernel (lbm_device.cuh), hence the same bits as two
// one-step launches (tests). Ghost rows of two-row slabs (gx = 2) supply the dependency cone across GPUs.
// -------------------------------------------------------------------------------------------------------
template <int T, bool HALO, bool PROBE>
__global__ void __launch_bounds__(T, 4) k_step2x(const __grid_constant__ StepParams P)
{
    extern __shared__ double ring[];   // [4 slots][6 populations: 2,4,5,6,7,8][2T columns]
    if (HALO && !halo_wait(P)) return;
    constexpr int W = 2 * T - 4, RS = 2 * T;
    const int tid = threadIdx.x;
    const int y0 = blockIdx.x * W;
    const int rb = blockIdx.y;         // segments of the first row range, then of the second one
    int x0, x1;
    {
        const int nsa = (P.na + P.seg - 1) / P.seg;
        if (rb < nsa) {
            x0 = P.row0a + rb * P.seg;
            x1 = min(x0 + P.seg, P.row0a + P.na);
        } else {
            x0 = P.row0b + (rb - nsa) * P.seg;
            x1 = min(x0 + P.seg, P.row0b + P.nb);
        }
    }
    int ca = y0 - 2 + 2 * tid;          // even column of this thread's intermediate pair (ca, ca + 1)
    ca = ca < 0 ? ca + P.NY : (ca >= P.NY ? ca - P.NY : ca);
    const int cm = ca == 0 ? P.NY - 1 : ca - 1;                  // left neighbour of the pair
    const int cq = ca + 2 >= P.NY ? ca + 2 - P.NY : ca + 2;      // right neighbour of the pair
    const long long pl = P.plane;
    const int yo = y0 + 2 * tid - 2;    // first output column of this thread
    const bool out_pair = tid >= 1 && tid <= T - 2 && yo < P.NY;

    auto wrapx = [&](int r) { return r < 0 ? r + P.NX : (r >= P.NX ? r - P.NX : r); };
    auto load = [&](int j, double (&ga)[9], double (&gb)[9]) {
        const double *r0 = P.src + (long long)wrapx(j) * P.pitch, *rm = P.src + (long long)wrapx(j - 1) * P.pitch,
                     *rp = P.src + (long long)wrapx(j + 1) * P.pitch;
        const double2 v0 = ld2(r0 + ca), v1 = ld2(rm + pl + ca), v3 = ld2(rp + 3 * pl + ca);
        ga[0] = v0.x; gb[0] = v0.y;
        ga[1] = v1.x; gb[1] = v1.y;
        ga[3] = v3.x; gb[3] = v3.y;
        ga[2] = ldS(r0 + 2 * pl + cm); gb[2] = ldS(r0 + 2 * pl + ca);
        ga[5] = ldS(rm + 5 * pl + cm); gb[5] = ldS(rm + 5 * pl + ca);
        ga[6] = ldS(rp + 6 * pl + cm); gb[6] = ldS(rp + 6 * pl + ca);
        ga[4] = ldS(r0 + 4 * pl + ca + 1); gb[4] = ldS(r0 + 4 * pl + cq);
        ga[7] = ldS(rp + 7 * pl + ca + 1); gb[7] = ldS(rp + 7 * pl + cq);
        ga[8] = ldS(rm + 8 * pl + ca + 1); gb[8] = ldS(rm + 8 * pl + cq);
    };
    const long long tc = PROBE ? *P.tc_in : 0;
    const int j0 = x0 - 1, j1 = x1;     // intermediate rows j0..j1 inclusive

    // L2 prefetch of the nine 2 KB source segments of row jj (cp.async.bulk.prefetch: no registers, no shared
    // memory, one thread per block); the real loads then hit L2.
    auto prefetch_row = [&](int jj) {
        const int c0 = max(y0 - 4, 0);
        const unsigned bytes = (unsigned)(min(y0 + 2 * T, P.pitch) - c0) * 8u;
        const double *r0 = P.src + (long long)wrapx(jj) * P.pitch + c0, *rm = P.src + (long long)wrapx(jj - 1) * P.pitch + c0,
                     *rp = P.src + (long long)wrapx(jj + 1) * P.pitch + c0;
        constexpr int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
#pragma unroll
        for (int i = 0; i < 9; i++)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((cx[i] == 1 ? rm : (cx[i] == -1 ? rp : r0)) + i * pl), "r"(bytes)
                         : "memory");
    };
    // first step: row j of S_{t+1} (this thread's pair of columns) from the pulled populations fa / fb
    auto first_step = [&](int j, const double (&fa)[9], const double (&fb)[9], double (&sa)[9], double (&sb)[9]) {
        double rho, ux, uy, p[9], e[9];
        const bool probe_row = PROBE && wrapx(j) == P.px;
        moments(fa, rho, ux, uy);
        if (probe_row && ca == P.py) {   // time t+1 (redundant rows/columns write identical values)
            double *slot = P.probe + 2 * ((tc + 1) % P.probe_cap);
            slot[0] = ux;
            slot[1] = uy;
        }
        eq_poly(ux, uy, p);
        eq_from_poly(rho, p, e);
        collide(fa, e, P.omega, sa);
        moments(fb, rho, ux, uy);
        if (probe_row && ca + 1 == P.py) {
            double *slot = P.probe + 2 * ((tc + 1) % P.probe_cap);
            slot[0] = ux;
            slot[1] = uy;
        }
        eq_poly(ux, uy, p);
        eq_from_poly(rho, p, e);
        collide(fb, e, P.omega, sb);
    };
    auto ring_store = [&](int j, const double (&sa)[9], const double (&sb)[9]) {
        double2 *slot = reinterpret_cast<double2 *>(ring + (size_t)(j & 3) * 6 * RS) + tid;
        slot[0 * T] = make_double2(sa[2], sb[2]);
        slot[1 * T] = make_double2(sa[4], sb[4]);
        slot[2 * T] = make_double2(sa[5], sb[5]);
        slot[3 * T] = make_double2(sa[6], sb[6]);
        slot[4 * T] = make_double2(sa[7], sb[7]);
        slot[5 * T] = make_double2(sa[8], sb[8]);
    };
    // second step: row r of S_{t+2} from intermediate rows r-1 (A), r (B), r+1 (D) in the ring and the unshifted
    // populations handed over in registers (0 of row r, 1 of row r-1, 3 of row r+1)
    auto second_step = [&](int r, double h0a, double h0b, double h1a, double h1b, double h3a, double h3b) {
        const double *A = ring + (size_t)((r - 1) & 3) * 6 * RS + 2 * tid;
        const double *B = ring + (size_t)(r & 3) * 6 * RS + 2 * tid;
        const double *D = ring + (size_t)((r + 1) & 3) * 6 * RS + 2 * tid;
        double ha[9], hb[9];
        ha[0] = h0a;           hb[0] = h0b;
        ha[1] = h1a;           hb[1] = h1b;
        ha[3] = h3a;           hb[3] = h3b;
        ha[2] = B[0 * RS - 1]; hb[2] = B[0 * RS];
        ha[4] = B[1 * RS + 1]; hb[4] = B[1 * RS + 2];
        ha[5] = A[2 * RS - 1]; hb[5] = A[2 * RS];
        ha[6] = D[3 * RS - 1]; hb[6] = D[3 * RS];
        ha[7] = D[4 * RS + 1]; hb[7] = D[4 * RS + 2];
        ha[8] = A[5 * RS + 1]; hb[8] = A[5 * RS + 2];
        const int xo = wrapx(r);
        const bool probe_row = PROBE && xo == P.px;
        double ta[9], tb[9];
        {
            double rho, ux, uy, p[9], e[9];
            moments(ha, rho, ux, uy);
            if (probe_row && yo == P.py) {        // time t+2: also advances the device clock
                double *slot = P.probe + 2 * ((tc + 2) % P.probe_cap);
                slot[0] = ux;
                slot[1] = uy;
                *P.tc_out = tc + 2;
                publish_progress(P, tc + 2);   // (this thread also wrote the sample of t+1, in first_step)
            }
            eq_poly(ux, uy, p);
            eq_from_poly(rho, p, e);
            collide(ha, e, P.omega, ta);
            moments(hb, rho, ux, uy);
            if (probe_row && yo + 1 == P.py) {
                double *slot = P.probe + 2 * ((tc + 2) % P.probe_cap);
                slot[0] = ux;
                slot[1] = uy;
                *P.tc_out = tc + 2;
                publish_progress(P, tc + 2);   // (this thread also wrote the sample of t+1, in first_step)
            }
            eq_poly(ux, uy, p);
            eq_from_poly(rho, p, e);
            collide(hb, e, P.omega, tb);
        }
        double *o = P.dst + (long long)xo * P.pitch + yo;
#pragma unroll
        for (int i = 0; i < 9; i++) st2(o + i * pl, ta[i], tb[i]);
        if (HALO) {   // no ghost snapshot: nothing is materialised from a two-step pass
            store_halo(P, xo, yo, ta);
            store_halo(P, xo, yo + 1, tb);
        }
    };

    double fa[9], fb[9];
    {
        // One row per trip: first step of row j -> ring -> barrier -> second step of row j-1. fa/fb are dead once
        // row j is collided, so the loads of row j+1 are issued right there, into the same registers, and fly while
        // the second step is computed. Unshifted populations in registers: 0 of row j-1, 1 of rows j-1 and j-2.
        double a0p = 0, b0p = 0, a1p = 0, b1p = 0, a1pp = 0, b1pp = 0;
        load(j0, fa, fb);
        for (int j = j0; j <= j1; j++) {
            if (tid == 0 && P.pf && j + P.pf <= j1) prefetch_row(j + P.pf);
            double sa[9], sb[9];
            first_step(j, fa, fb, sa, sb);
            if (j < j1) load(j + 1, fa, fb);
            ring_store(j, sa, sb);
            __syncthreads();
            if (j >= x0 + 1 && out_pair) second_step(j - 1, a0p, b0p, a1pp, b1pp, sa[3], sb[3]);
            a1pp = a1p;
            b1pp = b1p;
            a1p = sa[1];
            b1p = sb[1];
            a0p = sa[0];
            b0p = sb[0];
        }
    }
    if (HALO) halo_signal(P);
}

// -------------------------------------------------------------------------------------------------------
// D time steps per pass (D = 2, 3, 4): the temporal blocking of k_step2x carried further, so that a cell update
// moves 144 / D bytes (+ overlap) through DRAM. Same geometry — a block of T threads marches along x over a
// segment of rows, every thread owns an aligned pair of columns of every intermediate state — but D - 1 rings:
//   iteration j, phase A : level 1 = row j of S_{t+1} from global S_t (loads issued one row ahead) -> ring 0
//                __syncthreads (the only one per row)
//                phase B : level d = 2..D on row j - (2d-3): pulled from ring d-2, collided, stored to ring d-1
//                          (d < D) or to global S_{t+D} (d = D). Level 2 reads what phase A of this iteration wrote;
//                          every deeper level reads rows its ring received in EARLIER iterations, so the D - 1 levels
//                          of phase B are independent of each other (instruction-level parallelism) and need no
//                          barrier between them.
// Ring layout: a population written for intermediate row q is read when the consumer reaches row q+1 (5, 8), q (2, 4)
// or q-1 (6, 7), so 4 / 3 / 2 slots are enough: 18 slot-populations x 2T doubles = 36 KB per ring at T = 128 (24 in
// k_step2x), which is what lets three blocks of the three-step kernel share an SM. The populations that do not move
// along y (0, 1, 3) are handed from level to level in registers.
// Redundant work: 2(D-1) pairs of columns per strip and D-1 rows per level at each end of a segment.
// FINAL: the last level stops after the moments and writes reference-layout f_post / rho / u of time t+D — how results
// are materialised after a call that ended on a multi-step pass (the other buffer still holds S_t).
// -------------------------------------------------------------------------------------------------------
template <int T, int D>
struct Deep {
    static constexpr int W = 2 * T - 4 * (D - 1);   // output columns per block
    static constexpr int RS = 2 * T;                // doubles per ring row
    static constexpr int SP = 18;                   // slot-populations per ring
    static constexpr int MINB = D == 2 ? 3 : 2;     // resident blocks per SM the kernel is compiled for (D >= 3: 255 registers)
    static constexpr int SMEM = (D - 1) * SP * RS * (int)sizeof(double);
};

template <int T, int D, bool HALO, bool PROBE, bool FINAL>
__global__ void __launch_bounds__(T, Deep<T, D>::MINB) k_stepNx(const __grid_constant__ StepParams P)
{
    extern __shared__ double ring[];   // [D-1 rings][18 slot-populations][2T columns]
    static_assert(D >= 2 && D <= 4, "depth");
    if (HALO && !halo_wait(P)) return;
    constexpr int W = Deep<T, D>::W, RS = Deep<T, D>::RS, SP = Deep<T, D>::SP;
    constexpr bool LATE_LOAD = D >= 3;   // the next row's loads are issued before the LAST level of phase B only
    const int tid = threadIdx.x;
    const int y0 = (blockIdx.x + P.strip0) * W;
    int x0, x1;
    {
        const int rb = blockIdx.y;     // segments of the first row range, then of the second one
        const int nsa = (P.na + P.seg - 1) / P.seg;
        if (rb < nsa) {
            x0 = P.row0a + rb * P.seg;
            x1 = min(x0 + P.seg, P.row0a + P.na);
        } else {
            x0 = P.row0b + (rb - nsa) * P.seg;
            x1 = min(x0 + P.seg, P.row0b + P.nb);
        }
    }
    const int yo = y0 - 2 * (D - 1) + 2 * tid;     // (unwrapped) even column of this thread's pair, at every level
    const int ca = yo < 0 ? yo + P.NY : (yo >= P.NY ? yo - P.NY : yo);
    const int cm = ca == 0 ? P.NY - 1 : ca - 1;                  // left neighbour of the pair
    const int cq = ca + 2 >= P.NY ? ca + 2 - P.NY : ca + 2;      // right neighbour of the pair
    const long long pl = P.plane;

    auto wrapx = [&](int r) { return r < 0 ? r + P.NX : (r >= P.NX ? r - P.NX : r); };
    auto load = [&](int j, double (&ga)[9], double (&gb)[9]) {
        const double *r0 = P.src + (long long)wrapx(j) * P.pitch, *rm = P.src + (long long)wrapx(j - 1) * P.pitch,
                     *rp = P.src + (long long)wrapx(j + 1) * P.pitch;
        const double2 v0 = ld2(r0 + ca), v1 = ld2(rm + pl + ca), v3 = ld2(rp + 3 * pl + ca);
        ga[0] = v0.x; gb[0] = v0.y;
        ga[1] = v1.x; gb[1] = v1.y;
        ga[3] = v3.x; gb[3] = v3.y;
        ga[2] = ldS(r0 + 2 * pl + cm); gb[2] = ldS(r0 + 2 * pl + ca);
        ga[5] = ldS(rm + 5 * pl + cm); gb[5] = ldS(rm + 5 * pl + ca);
        ga[6] = ldS(rp + 6 * pl + cm); gb[6] = ldS(rp + 6 * pl + ca);
        ga[4] = ldS(r0 + 4 * pl + ca + 1); gb[4] = ldS(r0 + 4 * pl + cq);
        ga[7] = ldS(rp + 7 * pl + ca + 1); gb[7] = ldS(rp + 7 * pl + cq);
        ga[8] = ldS(rm + 8 * pl + ca + 1); gb[8] = ldS(rm + 8 * pl + cq);
    };
    auto prefetch_row = [&](int jj) {   // the nine source segments of level-1 row jj -> L2 (one thread per block)
        const int c0 = max(y0 - 2 * D, 0);
        const unsigned bytes = (unsigned)(min(y0 - 2 * (D - 1) + 2 * T + 2, P.pitch) - c0) * 8u;
        const double *r0 = P.src + (long long)wrapx(jj) * P.pitch + c0, *rm = P.src + (long long)wrapx(jj - 1) * P.pitch + c0,
                     *rp = P.src + (long long)wrapx(jj + 1) * P.pitch + c0;
        constexpr int cx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
#pragma unroll
        for (int i = 0; i < 9; i++)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((cx[i] == 1 ? rm : (cx[i] == -1 ? rp : r0)) + i * pl), "r"(bytes)
                         : "memory");
    };
    const long long tc = PROBE ? *P.tc_in : 0;
    auto probe_put = [&](int lvl, double ux, double uy) {   // sample of time t+lvl (redundant rows / columns of
        double *slot = P.probe + 2 * ((tc + lvl) % P.probe_cap);   // neighbouring blocks write identical values)
        slot[0] = ux;
        slot[1] = uy;
        if (lvl == D) {          // this thread also wrote the samples of t+1 .. t+D-1 (same pair, same row)
            *P.tc_out = tc + D;
            publish_progress(P, tc + D);
        }
    };
    // Ring rows hold the even columns of the block in [0, T) and the odd ones in [T, 2T): the +-1 column shifts of the
    // pulls are then unit-stride 64-bit accesses (no bank conflicts; pairs stored as one 128-bit word cost every
    // shifted load two extra wavefronts — 2.5e8 conflicts per launch in the first version, profiles/r02_summary.md).
    const int tm = tid > 0 ? tid - 1 : 0, tp = tid < T - 1 ? tid + 1 : T - 1;   // (clamped: edge pairs compute garbage nobody reads)
    auto ring_store = [&](int b, unsigned q, const double (&sa)[9], const double (&sb)[9]) {
        double *R = ring + (size_t)b * SP * RS + tid;
        const unsigned q3 = q % 3u, q4 = q & 3u, q2 = q & 1u;
        R[(0 + q3) * RS] = sa[2];   R[(0 + q3) * RS + T] = sb[2];
        R[(3 + q3) * RS] = sa[4];   R[(3 + q3) * RS + T] = sb[4];
        R[(6 + q4) * RS] = sa[5];   R[(6 + q4) * RS + T] = sb[5];
        R[(10 + q4) * RS] = sa[8];  R[(10 + q4) * RS + T] = sb[8];
        R[(14 + q2) * RS] = sa[6];  R[(14 + q2) * RS + T] = sb[6];
        R[(16 + q2) * RS] = sa[7];  R[(16 + q2) * RS + T] = sb[7];
    };
    // the six y-moving pulls of the pair on intermediate row q: 5, 8 from row q-1; 2, 4 from row q; 6, 7 from row q+1.
    // Even cell (column 2t): c_y = +1 pulls the odd column of pair t-1, c_y = -1 the odd column of pair t;
    // odd cell (column 2t+1): c_y = +1 pulls the even column of pair t, c_y = -1 the even column of pair t+1.
    auto ring_gather = [&](int b, unsigned q, double (&ha)[9], double (&hb)[9]) {
        const double *R = ring + (size_t)b * SP * RS;
        const unsigned m3 = q % 3u, a4 = (q - 1u) & 3u, d2 = (q + 1u) & 1u;
        ha[2] = R[(0 + m3) * RS + T + tm];   hb[2] = R[(0 + m3) * RS + tid];
        ha[4] = R[(3 + m3) * RS + T + tid];  hb[4] = R[(3 + m3) * RS + tp];
        ha[5] = R[(6 + a4) * RS + T + tm];   hb[5] = R[(6 + a4) * RS + tid];
        ha[8] = R[(10 + a4) * RS + T + tid]; hb[8] = R[(10 + a4) * RS + tp];
        ha[6] = R[(14 + d2) * RS + T + tm];  hb[6] = R[(14 + d2) * RS + tid];
        ha[7] = R[(16 + d2) * RS + T + tid]; hb[7] = R[(16 + d2) * RS + tp];
    };

    // unshifted populations in flight between the levels (a = even column, b = odd column of the pair):
    //   level 1 -> 2: population 0 of row j-1, population 1 of rows j-1 and j-2 (population 3 of row j is fresh)
    //   level d -> d+1 (d >= 2), w = the row level d wrote last: 3 of row w; 0 of rows w, w-1; 1 of rows w, w-1, w-2
    double g0a = 0, g0b = 0, g1a[2] = {0, 0}, g1b[2] = {0, 0};
    double k3a[D] = {}, k3b[D] = {}, k0a[D][2] = {}, k0b[D][2] = {}, k1a[D][3] = {}, k1b[D][3] = {};
    const int jbeg = x0 - (D - 1), jend = x1 - 1 + (2 * D - 3), l1end = x1 + (D - 1);
    double fa[9], fb[9];
    load(jbeg, fa, fb);
#pragma unroll 1
    for (int j = jbeg; j <= jend; j++) {
        const unsigned q = (unsigned)(j - jbeg) + 16u;   // ring row counter of level-1 row j
        // ---- phase A: level 1, row j (past the end of the segment: harmless recomputation of the last loaded row)
        double sa[9], sb[9];
        {
            if (tid == 0 && P.pf && j + P.pf < l1end) prefetch_row(j + P.pf);
            double uax, uay, ubx, uby;
            bool slow_a = false, slow_b = false;
            relax_fast(fa, P.omega, sa, uax, uay, slow_a);
            relax_fast(fb, P.omega, sb, ubx, uby, slow_b);
            if (slow_a | slow_b) {   // operands outside the fast paths' range: never in a physical run
                if (slow_a) relax_redo(fa, P.omega, sa, uax, uay);
                if (slow_b) relax_redo(fb, P.omega, sb, ubx, uby);
            }
            if (PROBE && j < l1end && wrapx(j) == P.px) {
                if (ca == P.py) probe_put(1, uax, uay);
                if (ca + 1 == P.py) probe_put(1, ubx, uby);
            }
            if (!LATE_LOAD && j + 1 < l1end) load(j + 1, fa, fb);
            ring_store(0, q, sa, sb);
        }
        __syncthreads();
        // ---- phase B: levels 2..D, all from ring rows that are complete; computed unconditionally (pipeline fill and
        // drain, edge pairs: garbage in, garbage out, nothing stored) so that they form ONE basic block
        double ta[D + 1][9], tb[D + 1][9], vax[D + 1], vay[D + 1], vbx[D + 1], vby[D + 1];
        double fin_a[9], fin_b[9];   // FINAL: the last level's pulled populations are the result
        bool slow = false, sl_a[D + 1] = {}, sl_b[D + 1] = {};
        auto gather_level = [&](int d, double (&ha)[9], double (&hb)[9]) {
            ring_gather(d - 2, q - (unsigned)(2 * d - 3), ha, hb);
            if (d == 2) {
                ha[0] = g0a;    hb[0] = g0b;
                ha[1] = g1a[1]; hb[1] = g1b[1];
                ha[3] = sa[3];  hb[3] = sb[3];
            } else {
                ha[0] = k0a[d - 2][1]; hb[0] = k0b[d - 2][1];
                ha[1] = k1a[d - 2][2]; hb[1] = k1b[d - 2][2];
                ha[3] = k3a[d - 2];    hb[3] = k3b[d - 2];
            }
        };
        auto redo_level = [&](int d) {   // operands outside the fast paths' range (never in a physical run): pull again, library arithmetic
            const double om = d < D ? P.omega : P.omega_last;
            double ha[9], hb[9];
            gather_level(d, ha, hb);
            if (sl_a[d]) relax_redo(ha, om, ta[d], vax[d], vay[d]);
            if (sl_b[d]) relax_redo(hb, om, tb[d], vbx[d], vby[d]);
        };
        auto store_level = [&](int d) {
            const int lag = 2 * d - 3, r = j - lag;
            const bool act = r >= x0 - (D - d) && r < x1 + (D - d);
            const bool mine = tid >= d - 1 && tid <= T - d && (d < D || yo < P.NY);
            if (act && mine) {
                const int xo = wrapx(r);
                if (PROBE && (d < D || !FINAL) && xo == P.px) {
                    if (ca == P.py) probe_put(d, vax[d], vay[d]);
                    if (ca + 1 == P.py) probe_put(d, vbx[d], vby[d]);
                }
                if (d < D) {
                    ring_store(d - 1, q - (unsigned)lag, ta[d], tb[d]);
                } else if (!FINAL) {
                    double *o = P.dst + (long long)xo * P.pitch + yo;
#pragma unroll
                    for (int i = 0; i < 9; i++) st2(o + i * pl, ta[d][i], tb[d][i]);
                    if (HALO) {   // (no ghost snapshot: results of a slab are materialised from its own rows only)
                        store_halo(P, xo, yo, ta[d]);
                        store_halo(P, xo, yo + 1, tb[d]);
                    }
                } else {
#pragma unroll
                    for (int c2 = 0; c2 < 2; c2++) {
                        const double(&h)[9] = c2 ? fin_b : fin_a;
                        const int y = yo + c2;
                        if (y < P.oy0 || y >= P.oy0 + P.ow) continue;
                        double rho, ux, uy;
                        moments(h, rho, ux, uy);
                        const long long o = (long long)(xo - P.ox0) * P.ow + (y - P.oy0);
                        if (P.o_f) {
#pragma unroll
                            for (int i = 0; i < 9; i++) P.o_f[o * 9 + i] = h[i];
                        }
                        if (P.o_rho) P.o_rho[o] = rho;
                        if (P.o_u) {
                            P.o_u[o * 2] = ux;
                            P.o_u[o * 2 + 1] = uy;
                        }
                    }
                }
            }
            if (d < D) {   // hand the unshifted populations of the row just written to level d+1
                k1a[d - 1][2] = k1a[d - 1][1]; k1a[d - 1][1] = k1a[d - 1][0]; k1a[d - 1][0] = ta[d][1];
                k1b[d - 1][2] = k1b[d - 1][1]; k1b[d - 1][1] = k1b[d - 1][0]; k1b[d - 1][0] = tb[d][1];
                k0a[d - 1][1] = k0a[d - 1][0]; k0a[d - 1][0] = ta[d][0];
                k0b[d - 1][1] = k0b[d - 1][0]; k0b[d - 1][0] = tb[d][0];
                k3a[d - 1] = ta[d][3];
                k3b[d - 1] = tb[d][3];
            }
        };
#pragma unroll
        for (int d = D; d >= 2; d--) {
            // the next row's loads fly during the LAST level of this phase only (its source segments are in L2
            // already): the levels before it compute without 36 registers of loads in flight
            if (LATE_LOAD && d == 2 && j + 1 < l1end) load(j + 1, fa, fb);
            double ha[9], hb[9];
            gather_level(d, ha, hb);
            sl_a[d] = sl_b[d] = false;
            if (d < D || !FINAL) {
                const double om = d < D ? P.omega : P.omega_last;
                relax_fast(ha, om, ta[d], vax[d], vay[d], sl_a[d]);
                relax_fast(hb, om, tb[d], vbx[d], vby[d], sl_b[d]);
                slow |= sl_a[d] | sl_b[d];
            } else {
#pragma unroll
                for (int i = 0; i < 9; i++) {
                    fin_a[i] = ha[i];
                    fin_b[i] = hb[i];
                }
            }
        }
        // all levels of phase B form one basic block up to here (four cells in flight for D = 3); the cold redo and the stores follow
        if (slow) {
#pragma unroll
            for (int d = D; d >= 2; d--) redo_level(d);
        }
#pragma unroll
        for (int d = D; d >= 2; d--) store_level(d);
        g1a[1] = g1a[0]; g1a[0] = sa[1]; g0a = sa[0];
        g1b[1] = g1b[0]; g1b[0] = sb[1]; g0b = sb[0];
    }
    if (HALO) halo_signal(P);
}

// ==== HOT KERNELS END ====

// -------------------------------------------------------------------------------------------------------
// Launch-bound lattices (BASELINE.json configs 1-3: 5 000 - 10 000 cells, 2 500 - 40 000 steps per run): MANY time
// steps in ONE launch of a single thread-block cluster. The whole lattice lives in the cluster's distributed shared
// memory — CTA r owns rows [r NX / C, (r+1) NX / C) of both A/B buffers — a thread owns one cell (or two) for the whole launch,
// so its kind, its three source-row pointers (local or a neighbour CTA's shared memory) and its destination are
// computed once; a step is nine shared-memory pulls, the common cell update, nine stores and one hardware cluster
// barrier (~0.2 us) instead of a kernel boundary (~2 us inside a replayed graph). The last two states are written
// back to the global A/B buffers at the end, so everything else (materialisation, omega redo, probe clock) is as
// after a one-step launch. Same per-cell arithmetic (lbm_device.cuh), same rule table: same bits.
// (A grid-wide cooperative version with a software barrier lost to graph replay in round 1,
//  profiles/r01e_persistent_vs_graph.txt; the cluster barrier is what changes the balance.)
// -------------------------------------------------------------------------------------------------------
struct ClusterParams {
    StepParams S;      // src = S_t (global), dst = the other global buffer; NX, NY, pitch, plane, tables, probe, omega
    int n_steps;
    int R;             // rows per CTA
    int n_cells;       // R * NY
    double *ob[2];     // outlet side buffers of the source / the other global buffer
};

// Shared-memory addresses inside the cluster's window (32 bits): one `mapa` per source row at kernel start, then a
// population of ANY cell of the cluster is one `ld.shared::cluster.f64` — own CTA or a neighbour's.
__device__ __forceinline__ unsigned dsmem_map(unsigned addr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ double dsmem_ld(unsigned addr)
{
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

template <bool MASK, int M, int TMAX>
__global__ void __launch_bounds__(TMAX, 1) k_cluster_steps(const __grid_constant__ ClusterParams Q)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ double sm[];   // [2 buffers][9][most rows per CTA][NY]
    const StepParams &P = Q.S;
    const int rank = (int)cluster.block_rank(), C = (int)cluster.num_blocks();
    const int NY = P.NY, NX = P.NX, plane = Q.n_cells;
    const int bstride = 9 * plane;                 // doubles per buffer
    // balanced split: CTA r owns rows [r NX / C, (r+1) NX / C) — sizes differ by at most one row, so any NX >= C uses
    // all C SMs (100 rows on 16 CTAs: 6 or 7 rows each)
    auto lo = [&](int r) { return (int)(((long long)r * NX) / C); };
    auto own = [&](int x) { return (int)(((long long)(x + 1) * C - 1) / NX); };
    const int row_lo = lo(rank), nrows = lo(rank + 1) - row_lo;
    const int tid = threadIdx.x, T = blockDim.x;
    constexpr int kcx[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}, kcy[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
    constexpr int kopp[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6};
    constexpr int G = (M == 2 && TMAX <= 512) ? 2 : 1;

    // S_t of my rows: global -> buffer 0
    for (int q = tid; q < nrows * NY; q += T) {
        const int r = q / NY, y = q - r * NY;
        const double *g = P.src + (long long)(row_lo + r) * P.pitch + y;
#pragma unroll
        for (int i = 0; i < 9; i++) sm[i * plane + q] = __ldcg(g + i * P.plane);
    }
    // my cell(s): everything that does not change from step to step. Where population i comes from is ONE address per
    // population for every kind of cell (the streamed neighbour, the cell's own opposite population for a bounce-back
    // rule, the cell itself where the value is replaced afterwards), so the nine loads of a step are the same straight
    // code for fluid and boundary cells; a boundary cell then patches its populations (wall velocity term, inlet
    // constant, outlet copy) in a short branch.
    bool act[M];
    int cy_[M], cq[M], cx_[M];
    unsigned kind[M], src[M][9], todo[M];   // todo: bit i = subtract the wall term, bit 9+i = inlet constant, bit 18+i = outlet copy
    lbm_kind kd[M];
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(sm);
#pragma unroll
    for (int m = 0; m < M; m++) {
        const int q = tid + m * T;
        act[m] = q < nrows * NY;
        const int r = act[m] ? q / NY : 0, y = act[m] ? q - r * NY : 0, x = row_lo + r;
        cq[m] = act[m] ? q : 0;
        cx_[m] = x;
        cy_[m] = y;
        const int ym = y == 0 ? NY - 1 : y - 1, yp = y == NY - 1 ? 0 : y + 1;
        const int xm = x == 0 ? NX - 1 : x - 1, xp = x == NX - 1 ? 0 : x + 1;
        const int om = own(xm), op = own(xp);
        const unsigned b0 = dsmem_map(sm0, rank) + 8u * (unsigned)(r * NY);
        const unsigned bm = dsmem_map(sm0, om) + 8u * (unsigned)((xm - lo(om)) * NY);
        const unsigned bp = dsmem_map(sm0, op) + 8u * (unsigned)((xp - lo(op)) * NY);
        kind[m] = 0;
        todo[m] = 0;
        kd[m] = lbm_kind{};
        if (MASK && act[m]) {
            kind[m] = P.kind_map[(long long)x * P.pitch + y];
            if (kind[m]) kd[m] = P.kinds[kind[m]];
        }
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const unsigned row = kcx[i] == 1 ? bm : (kcx[i] == -1 ? bp : b0);
            const int col = kcy[i] == 1 ? ym : (kcy[i] == -1 ? yp : y);
            unsigned a = row + 8u * (unsigned)(i * plane + col);
            if (MASK && kind[m]) {
                const int type = kd[m].rule[i] & 7, trow = kd[m].rule[i] >> 3;
                if (type == LBM_RULE_BOUNCE) {
                    a = b0 + 8u * (unsigned)(kopp[i] * plane + y);
                    if (trow) todo[m] |= 1u << i;
                } else if (type != LBM_RULE_PULL) {
                    a = b0 + 8u * (unsigned)(i * plane + y);
                    todo[m] |= 1u << ((type == LBM_RULE_CONST ? 9 : 18) + i);
                }
            }
            src[m][i] = a;
        }
    }
    const long long tc0 = P.probe ? *P.tc_in : 0;
    cluster.sync();

    for (int s = 0; s < Q.n_steps; s++) {
        const unsigned so = (s & 1) ? 8u * (unsigned)bstride : 0u;
        const int dofs = (s & 1) ? 0 : bstride;
        // two cells of a thread are interleaved (G = 2) where the block's register budget allows it (<= 512 threads: 128
        // registers), else they are updated one after the other
#pragma unroll
        for (int g = 0; g < M; g += G) {
            double f[M][9];
#pragma unroll
            for (int m = g; m < g + G; m++)
#pragma unroll
                for (int i = 0; i < 9; i++) f[m][i] = dsmem_ld(src[m][i] + so);
            if (MASK) {
#pragma unroll
                for (int m = g; m < g + G; m++) {
                    const unsigned td = todo[m];
                    if (!td) continue;               // fluid cells and plain bounce-back (rigid walls): nothing to patch
                    const double *oc = Q.ob[s & 1];
#pragma unroll
                    for (int i = 1; i < 9; i++)
                        if ((td >> i) & 1) f[m][i] = sub(f[m][i], P.ktab[(kd[m].rule[i] >> 3) * 9 + kopp[i]]);
                    if (td >> 9) {
#pragma unroll
                        for (int i = 0; i < 9; i++)
                            if ((td >> (9 + i)) & 1) f[m][i] = P.ctab[(kd[m].rule[i] >> 3) * 9 + i];
                        if ((td >> 21) & 1) f[m][3] = __ldcg(oc + 0 * P.pitch + cy_[m]);
                        if ((td >> 24) & 1) f[m][6] = __ldcg(oc + 1 * P.pitch + cy_[m]);
                        if ((td >> 25) & 1) f[m][7] = __ldcg(oc + 2 * P.pitch + cy_[m]);
                    }
                }
            }
            // branch-free division / square root (lbm_device.cuh): the dependent chain of ONE cell update is what a step of
            // this kernel waits for, and two cells of a thread interleave; operands the fast paths reject redo the moments
            // with the library operations
            double rho[M], ux[M], uy[M], p[M][9];
            bool slow[M];
#pragma unroll
            for (int m = g; m < g + G; m++) {
                slow[m] = false;
                moments_fast(f[m], rho[m], ux[m], uy[m], slow[m]);
                eq_poly_fast(ux[m], uy[m], p[m], slow[m]);
            }
#pragma unroll
            for (int m = g; m < g + G; m++) {
                if (slow[m]) {
                    moments(f[m], rho[m], ux[m], uy[m]);
                    eq_poly(ux[m], uy[m], p[m]);
                }
            }
#pragma unroll
            for (int m = g; m < g + G; m++) {
                if (!act[m]) continue;
                const int y = cy_[m];
                if (P.probe && cx_[m] == P.px && y == P.py) {
                    const long long t_new = tc0 + s + 1;
                    double *slot = P.probe + 2 * (t_new % P.probe_cap);
                    slot[0] = ux[m];
                    slot[1] = uy[m];
                    // the time word the host polls: every 32nd step only — its system-scope fence holds up this
                    // thread, hence (cluster barrier) every CTA, for ~1 us; the samples of a finished launch need no word
                    if ((s & 31) == 31) publish_progress(P, t_new);
                }
                double e[9], o[9];
                eq_from_poly(rho[m], p[m], e);
                collide(f[m], e, P.omega, o);
                const unsigned flags = MASK ? kd[m].flags : 0u, skip = MASK ? kd[m].skip_store : 0u;
                if (MASK && (flags & LBM_CELL_OUTLET_SRC)) {
                    double *on = Q.ob[(s + 1) & 1];
                    __stcg(on + 0 * P.pitch + y, f[m][3]);
                    __stcg(on + 1 * P.pitch + y, f[m][6]);
                    __stcg(on + 2 * P.pitch + y, f[m][7]);
                    __threadfence();
                }
                double *w = sm + dofs + cq[m];
                if (!MASK || skip == 0) {
#pragma unroll
                    for (int i = 0; i < 9; i++) w[i * plane] = o[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 9; i++)
                        if (!((skip >> i) & 1)) w[i * plane] = o[i];
                }
                if (MASK && (flags & (LBM_CELL_PBC_IN_SRC | LBM_CELL_PBC_OUT_SRC))) {
                    // periodic_with_pressure_variations (boundary_conditions.py:337-344), see store_pbc: the virtual rows 0 and
                    // NX-1 belong to the first / last CTA of the cluster
                    if (flags & LBM_CELL_PBC_IN_SRC) {
                        const double w1 = mul(LBM_W1, P.rho_in), w5 = mul(LBM_W5, P.rho_in);
                        double *v = cluster.map_shared_rank(sm, 0) + dofs + y;   // row 0
                        v[1 * plane] = add(mul(w1, p[m][1]), sub(o[1], e[1]));
                        v[5 * plane] = add(mul(w5, p[m][5]), sub(o[5], e[5]));
                        v[8 * plane] = add(mul(w5, p[m][8]), sub(o[8], e[8]));
                    }
                    if (flags & LBM_CELL_PBC_OUT_SRC) {
                        const double w1 = mul(LBM_W1, P.rho_out), w5 = mul(LBM_W5, P.rho_out);
                        double *v = cluster.map_shared_rank(sm, C - 1) + dofs + (NX - 1 - lo(C - 1)) * NY + y;   // row NX-1
                        v[3 * plane] = add(mul(w1, p[m][3]), sub(o[3], e[3]));
                        v[6 * plane] = add(mul(w5, p[m][6]), sub(o[6], e[6]));
                        v[7 * plane] = add(mul(w5, p[m][7]), sub(o[7], e[7]));
                    }
                }
            }
        }
        cluster.sync();
    }
    // S_{t+n} and S_{t+n-1} back to the global A/B buffers: the newest goes where n one-step launches would have left it
    const int n = Q.n_steps, nb = n & 1;
    double *g_new = (n & 1) ? P.dst : const_cast<double *>(P.src), *g_old = (n & 1) ? const_cast<double *>(P.src) : P.dst;
    for (int q = tid; q < nrows * NY; q += T) {
        const int r = q / NY, y = q - r * NY;
        const long long go = (long long)(row_lo + r) * P.pitch + y;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            __stcg(g_new + i * P.plane + go, sm[nb * bstride + i * plane + q]);
            __stcg(g_old + i * P.plane + go, sm[(nb ^ 1) * bstride + i * plane + q]);
        }
    }
    if (P.probe && rank == 0 && tid == 0) {   // device clocks of the two global buffers (tc_in: source buffer, tc_out: the other)
        long long *t_src = const_cast<long long *>(P.tc_in), *t_dst = P.tc_out;
        *((n & 1) ? t_dst : t_src) = tc0 + n;
        *((n & 1) ? t_src : t_dst) = tc0 + n - 1;
    }
}

// -------------------------------------------------------------------------------------------------------
// first collision of an uploaded / initialised state: S_0 = f + (feq(rho,u) - f)*omega with the GIVEN moments
// (lattice_boltzmann_method.py:213-215). Input is either reference-layout staging (rows [x0, x0+nrows)) or the
// separable initial fields of initial_values.py.
// -------------------------------------------------------------------------------------------------------
struct InitParams {
    StepParams S;
    const double *in_f, *in_rho, *in_u;   // AoS staging of rows [x0, x0+nrows), or null
    const double *rho_x, *ux_y;           // separable profiles (device), may be null
    double rho0, ux0, uy0;
    int x0, nrows;
};

template <bool HALO>
__global__ void __launch_bounds__(256) k_first_collide(const __grid_constant__ InitParams Q)
{
    const StepParams &P = Q.S;
    const int rb = blockIdx.x / P.bpr, cb = blockIdx.x - rb * P.bpr;
    const int x = Q.x0 + rb, y = cb * blockDim.x + threadIdx.x;
    if (y >= P.NY) return;
    double f[9], rho, ux, uy, p[9], e[9], s[9];
    if (Q.in_f) {
        const long long c = (long long)rb * P.NY + y;
#pragma unroll
        for (int i = 0; i < 9; i++) f[i] = Q.in_f[c * 9 + i];
        rho = Q.in_rho[c];
        ux = Q.in_u[2 * c];
        uy = Q.in_u[2 * c + 1];
        eq_poly(ux, uy, p);
        eq_from_poly(rho, p, e);
    } else {
        rho = Q.rho_x ? Q.rho_x[x] : Q.rho0;
        ux = Q.ux_y ? Q.ux_y[y] : Q.ux0;
        uy = Q.uy0;
        eq_poly(ux, uy, p);
        eq_from_poly(rho, p, e);
#pragma unroll
        for (int i = 0; i < 9; i++) f[i] = e[i];   // f = equilibrium_distr_func(density, velocity), experiments.py:122
    }
    collide(f, e, P.omega, s);
    unsigned flags = 0, skip = 0;
    if (P.kind_map) {
        const unsigned kind = P.kind_map[(long long)x * P.pitch + y];
        if (kind) {
            flags = P.kinds[kind].flags;
            skip = P.kinds[kind].skip_store;
        }
    }
    if (flags & LBM_CELL_OUTLET_SRC) {   // f_previous of the first step is the uploaded f itself
        P.out_next[0 * P.pitch + y] = f[3];
        P.out_next[1 * P.pitch + y] = f[6];
        P.out_next[2 * P.pitch + y] = f[7];
    }
    // ghost cells are owned by the neighbour that mirrors them (communicate() overwrites them before streaming)
    const bool ghost = x < P.gx || x >= P.NX - P.gx || (P.gy && (y == 0 || y == P.NY - 1));
    if (!ghost) {
        store_cell(P, x, y, s, skip);
        if (flags & (LBM_CELL_PBC_IN_SRC | LBM_CELL_PBC_OUT_SRC)) store_pbc(P, flags, y, s, p, e);
        if (HALO) store_halo(P, x, y, s);
    }
}

// -------------------------------------------------------------------------------------------------------
// stateless operators on reference-layout arrays
// -------------------------------------------------------------------------------------------------------
__global__ void k_equilibrium(long long n, const double *rho, const double *u, double *out)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double p[9], e[9];
    eq_poly(u[2 * c], u[2 * c + 1], p);
    eq_from_poly(rho[c], p, e);
#pragma unroll
    for (int i = 0; i < 9; i++) out[9 * c + i] = e[i];
}

__global__ void k_moments(long long n, const double *f, const double *rho_in, double *rho_out, double *u_out)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double g[9], rho, ux, uy;
#pragma unroll
    for (int i = 0; i < 9; i++) g[i] = f[9 * c + i];
    moments(g, rho, ux, uy);
    if (rho_out) rho_out[c] = rho;
    if (u_out) {
        if (rho_in) {   // compute_velocity_field(density, f) divides by the GIVEN density
            const double r = rho_in[c];
            const double jx = sub(add(add(g[1], g[5]), g[8]), add(add(g[3], g[6]), g[7]));
            const double jy = sub(add(add(g[2], g[5]), g[6]), add(add(g[4], g[7]), g[8]));
            ux = r != 0.0 ? div_rn(jx, r) : 0.0;
            uy = r != 0.0 ? div_rn(jy, r) : 0.0;
        }
        u_out[2 * c] = ux;
        u_out[2 * c + 1] = uy;
    }
}

__global__ void k_streaming_aos(int nx, int ny, const double *f, double *out)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)nx * ny) return;
    const int x = (int)(c / ny), y = (int)(c - (long long)x * ny);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        int xs = x - kCx[i], ys = y - kCy[i];
        xs = xs < 0 ? nx - 1 : (xs >= nx ? 0 : xs);
        ys = ys < 0 ? ny - 1 : (ys >= ny ? 0 : ys);
        out[9 * c + i] = f[((long long)xs * ny + ys) * 9 + i];
    }
}

__global__ void k_bc_apply_aos(int nx, int ny, const uint8_t *kind_map, const lbm_kind *kinds, const double *ktab,
                               const double *ctab, const double *f_pre, double *f_post, const double *f_prev)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)nx * ny) return;
    const unsigned kind = kind_map[c];
    if (!kind) return;
    const lbm_kind k = kinds[kind];
    const int x = (int)(c / ny);
    for (int i = 0; i < 9; i++) {
        const int type = k.rule[i] & 7, row = k.rule[i] >> 3;
        if (type == LBM_RULE_BOUNCE) {
            const int d = kOpp[i];
            double v = f_pre[9 * c + d];
            if (row) v = sub(v, ktab[row * 9 + d]);
            f_post[9 * c + i] = v;
        } else if (type == LBM_RULE_CONST) {
            f_post[9 * c + i] = ctab[row * 9 + i];
        } else if (type == LBM_RULE_OUTLET) {
            if (x > 0) f_post[9 * c + i] = f_prev[9 * (c - ny) + i];
        }
    }
}

__global__ void k_pbc_apply_aos(int nx, int ny, double rho_in, double rho_out, const double *rho, const double *u,
                                double *f_pre)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= ny) return;
    double p[9], e[9];
    {   // inflow: row 0 from row -2, populations 1,5,8
        const long long s = (long long)(nx - 2) * ny + y, d = y;
        eq_poly(u[2 * s], u[2 * s + 1], p);
        eq_from_poly(rho[s], p, e);
        const double w1 = mul(LBM_W1, rho_in), w5 = mul(LBM_W5, rho_in);
        const double v1 = add(mul(w1, p[1]), sub(f_pre[9 * s + 1], e[1]));
        const double v5 = add(mul(w5, p[5]), sub(f_pre[9 * s + 5], e[5]));
        const double v8 = add(mul(w5, p[8]), sub(f_pre[9 * s + 8], e[8]));
        f_pre[9 * d + 1] = v1;
        f_pre[9 * d + 5] = v5;
        f_pre[9 * d + 8] = v8;
    }
    __syncthreads();   // nx == 3 would alias; rows are distinct otherwise and each thread owns its y
    {   // outflow: row -1 from row 1, populations 3,6,7
        const long long s = (long long)1 * ny + y, d = (long long)(nx - 1) * ny + y;
        eq_poly(u[2 * s], u[2 * s + 1], p);
        eq_from_poly(rho[s], p, e);
        const double w1 = mul(LBM_W1, rho_out), w5 = mul(LBM_W5, rho_out);
        f_pre[9 * d + 3] = add(mul(w1, p[3]), sub(f_pre[9 * s + 3], e[3]));
        f_pre[9 * d + 6] = add(mul(w5, p[6]), sub(f_pre[9 * s + 6], e[6]));
        f_pre[9 * d + 7] = add(mul(w5, p[7]), sub(f_pre[9 * s + 7], e[7]));
    }
}

// min/max over packed rho / u staging (order-preserving integer image of a double)
__device__ __forceinline__ long long ord(double v)
{
    long long b = __double_as_longlong(v);
    return b < 0 ? (long long)(0x8000000000000000ULL - (unsigned long long)b) : b;
}
__host__ __device__ inline double unord(long long o)
{
    unsigned long long b = o < 0 ? (0x8000000000000000ULL - (unsigned long long)o) : (unsigned long long)o;
    double v;
    memcpy(&v, &b, 8);
    return v;
}

__global__ void k_minmax(long long n, const double *rho, const double *u, long long *acc /* [4] */)
{
    long long mn_r = 0x7fffffffffffffffLL, mx_r = -0x7fffffffffffffffLL - 1, mn_u = mn_r, mx_u = mx_r;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const long long r = ord(rho[c]), a = ord(u[2 * c]), b = ord(u[2 * c + 1]);
        mn_r = min(mn_r, r);
        mx_r = max(mx_r, r);
        mn_u = min(mn_u, min(a, b));
        mx_u = max(mx_u, max(a, b));
    }
    for (int o = 16; o; o >>= 1) {
        mn_r = min(mn_r, __shfl_xor_sync(0xffffffffu, mn_r, o));
        mx_r = max(mx_r, __shfl_xor_sync(0xffffffffu, mx_r, o));
        mn_u = min(mn_u, __shfl_xor_sync(0xffffffffu, mn_u, o));
        mx_u = max(mx_u, __shfl_xor_sync(0xffffffffu, mx_u, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(acc + 0, mn_r);
        atomicMax(acc + 1, mx_r);
        atomicMin(acc + 2, mn_u);
        atomicMax(acc + 3, mx_u);
    }
}

// -------------------------------------------------------------------------------------------------------
// Self-test of the hand-expanded arithmetic of lbm_device.cuh against the compiler's own IEEE operations:
// div_by(a, b, rcp_refined(b)) == __ddiv_rn(a, b) and sqrt_rn(a) == __dsqrt_rn(a), bit for bit (NaN == NaN),
// over operand classes chosen to reach every branch: raw random bit patterns (NaN, infinities, subnormals), the
// physical range (rho ~ 1, |j| < 0.2), numerators and quotients around the two range-test thresholds, signed
// zeros, exact quotients. b == 0 is outside div_by's contract (the callers test rho != 0) and is skipped.
// -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double bits(unsigned long long sign, unsigned long long expo, unsigned long long mant)
{
    return __longlong_as_double((long long)(((sign & 1) << 63) | ((expo & 0x7ff) << 52) | (mant & 0xfffffffffffffULL)));
}

__global__ void k_selftest_arith(long long n, unsigned long long seed, unsigned long long *out /* [2 counts][3 operands][2 fast-path counts] */)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long h0 = mix64(seed ^ (unsigned long long)i), h1 = mix64(h0), h2 = mix64(h1), h3 = mix64(h2);
    const double u1 = (double)(h2 >> 11) * 0x1p-53, u2 = (double)(h3 >> 11) * 0x1p-53;
    const double phys_b = 0.5 + 1.5 * u1, phys_a = (u2 - 0.5) * 0.4;
    double a, b;
    switch (i & 7) {
    case 0: a = __longlong_as_double((long long)h0); b = __longlong_as_double((long long)h1); break;
    case 1: a = phys_a; b = phys_b; break;
    case 2: a = bits(h0, 46 + (h0 >> 8) % 18, h1); b = (h0 & 2) ? phys_b : bits(h1 >> 60, 1023 - 40 + (h1 >> 40) % 80, h2); break;   // |a.hi| around 0x036
    case 3: a = bits(h0, h0 >> 1, h2); b = bits(h1, h1 >> 1, h3); break;                          // any exponents: tiny / huge quotients
    case 4: a = (h0 & 1) ? -0.0 : 0.0; b = (h0 & 2) ? __longlong_as_double((long long)h1) : ((h0 & 4) ? -phys_b : phys_b); break;
    case 5: a = phys_a; b = __longlong_as_double((long long)h1); break;
    case 6: a = __longlong_as_double((long long)h0); b = phys_b; break;
    default: b = phys_b; a = b * (double)((long long)(h0 % 2001) - 1000); break;                  // (nearly) exact quotients
    }
    if (b != 0.0) {
        const double mine = div_by(a, b, rcp_refined(b)), ref = __ddiv_rn(a, b);
        const bool same = __double_as_longlong(mine) == __double_as_longlong(ref) || (mine != mine && ref != ref);
        if (!same && atomicAdd(out + 0, 1ULL) == 0) {
            out[2] = (unsigned long long)__double_as_longlong(a);
            out[3] = (unsigned long long)__double_as_longlong(b);
        }
    }
    {
        const double mine = sqrt_rn(a), ref = __dsqrt_rn(a);
        const bool same = __double_as_longlong(mine) == __double_as_longlong(ref) || (mine != mine && ref != ref);
        if (!same && atomicAdd(out + 1, 1ULL) == 0) out[4] = (unsigned long long)__double_as_longlong(a);
    }
    // the branch-free variants of the multi-step kernels: whatever they do not flag as `slow` must be the IEEE result
    if (b != 0.0) {
        bool slow = false;
        const double mine = div_fast(a, b, rcp_refined(b), slow), ref = __ddiv_rn(a, b);
        const bool same = __double_as_longlong(mine) == __double_as_longlong(ref) || (mine != mine && ref != ref);
        if (!slow && !same && atomicAdd(out + 0, 1ULL) == 0) {
            out[2] = (unsigned long long)__double_as_longlong(a);
            out[3] = (unsigned long long)__double_as_longlong(b);
        }
        if (!slow) atomicAdd(out + 5, 1ULL);   // how many operand pairs took the fast path
    }
    {
        bool slow = false;
        const double mine = sqrt_fast(a, slow), ref = __dsqrt_rn(a);
        const bool same = __double_as_longlong(mine) == __double_as_longlong(ref) || (mine != mine && ref != ref);
        if (!slow && !same && atomicAdd(out + 1, 1ULL) == 0) out[4] = (unsigned long long)__double_as_longlong(a);
        if (!slow) atomicAdd(out + 6, 1ULL);
    }
}

