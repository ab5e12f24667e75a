/*
 * lbm_b200.h — C-ABI of the B200-native D2Q9 fp64 lattice Boltzmann time step.
 *
 * Drop-in boundary for ONE hot path of SimonSchrodi/lattice_boltzmann_parallel_solver:
 * `lattice_boltzmann_step` (src/lattice_boltzmann_method.py:191-228) with its boundary closures
 * (src/boundary_conditions.py:78-348, bundles src/boundary_utils.py:9-205) and halo exchange
 * (src/parallelization_utils.py:6-52). The reference is pure Python; what a maintainer binds is a ctypes
 * stub. Plain C types only: the caller owns every host pointer, the context owns every
 * device pointer. All functions return 0 on success or an LBM_ERR_* code; lbm_last_error() gives the text.
 * One context per process and GPU; a context is not thread-safe. Calls are stream-ordered on the context's
 * own streams and return without synchronising unless stated.
 *
 * Host array layout is the reference's: f[x][y][9] (C-contiguous, "AoS"), rho[x][y], u[x][y][2], float64.
 * Device layout is SoA: S[i][x][y] with the reference's axis 1 ("y") as the coalesced axis, rows padded to
 * 128 B. Device state is the POST-collision population set.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_OK 0
#define LBM_ERR_ARG 1      /* bad argument (AssertionError on the Python side, as the reference's asserts) */
#define LBM_ERR_CUDA 2     /* CUDA runtime failure */
#define LBM_ERR_STATE 3    /* call not valid in the context's current state */
#define LBM_ERR_NOMEM 4
#define LBM_ERR_TIMEOUT 5  /* a halo flag wait timed out (peer rank not stepping in lockstep) */

const char *lbm_last_error(void);
/* library version / build string (contains the compiled SM arch) */
const char *lbm_version(void);
/* number of visible CUDA devices, or a negative LBM_ERR_* */
int lbm_device_count(void);

/* ---------------------------------------------------------------------------------------------------------
 * Stateless operators on host arrays. Replace, bit for bit:
 *   lbm_equilibrium  <- equilibrium_distr_func   src/lattice_boltzmann_method.py:162-188
 *   lbm_density      <- compute_density          src/lattice_boltzmann_method.py:93-105
 *   lbm_velocity     <- compute_velocity_field   src/lattice_boltzmann_method.py:108-137
 *   lbm_streaming    <- streaming                src/lattice_boltzmann_method.py:140-159
 * n_cells = product of the leading dims. Each call copies in, runs one kernel, copies out, synchronises.
 * ------------------------------------------------------------------------------------------------------- */
int lbm_equilibrium(int device, int64_t n_cells, const double *rho, const double *u, double *f_out);
int lbm_density(int device, int64_t n_cells, const double *f, double *rho_out);
int lbm_velocity(int device, int64_t n_cells, const double *rho, const double *f, double *u_out);
int lbm_streaming(int device, int nx, int ny, const double *f, double *f_out);

/* ---------------------------------------------------------------------------------------------------------
 * Boundary description. The reference composes Python closures that overwrite entries of f_post one after
 * another (src/boundary_utils.py:52-53, 103-106, 164-203). Here the same effect is data: every lattice cell
 * carries a one-byte KIND; a kind says, per population i, where f_post[i] of that cell comes from.
 * ------------------------------------------------------------------------------------------------------- */
enum {
    LBM_RULE_PULL = 0,    /* f_post[i] = S[(x - c_i) mod n, i]           streaming,  lattice_boltzmann_method.py:153-157 */
    LBM_RULE_BOUNCE = 1,  /* f_post[i] = S[x, opp(i)] - K[row][opp(i)]   rigid_wall / moving_wall / rigid_object,
                                                                         boundary_conditions.py:108-109, 207-210, 153-163 */
    LBM_RULE_CONST = 2,   /* f_post[i] = C[row][i]                       inlet,      boundary_conditions.py:250-251 */
    LBM_RULE_OUTLET = 3   /* f_post[i] = f_prev[x-1, y, i]               outlet,     boundary_conditions.py:279-280 */
};
#define LBM_RULE(type, row) ((uint8_t)((type) | ((row) << 3)))   /* row: 0..31, row 0 of K is all zeros */

enum {
    LBM_CELL_OUTLET_SRC = 1,  /* this cell's f_post[3,6,7] is next step's f_prev[-2] (boundary_conditions.py:279-280) */
    LBM_CELL_PBC_IN_SRC = 2,  /* row -2: owns S[0, y, (1,5,8)] of the next step (boundary_conditions.py:339-340) */
    LBM_CELL_PBC_OUT_SRC = 4  /* row  1: owns S[-1, y, (3,6,7)]                 (boundary_conditions.py:343-344) */
};

typedef struct {
    uint8_t rule[9];      /* LBM_RULE(type,row) per population */
    uint8_t flags;        /* LBM_CELL_* */
    uint16_t skip_store;  /* bit i set: S'[i] of this cell is written by a PBC source cell instead */
} lbm_kind;

typedef struct {
    int n_kinds;               /* <= 256; kind 0 must be the all-PULL fluid cell */
    const lbm_kind *kinds;
    int n_k_rows;              /* <= 32 rows of 9 doubles; row 0 must be zeros */
    const double *k_table;     /* K_d of moving_wall, indexed [row][d] (d = the population that hits the wall) */
    int n_c_rows;              /* <= 32 rows of 9 doubles */
    const double *c_table;     /* inlet constants, [row][i] */
    double pbc_rho_in, pbc_rho_out;   /* p/c_s**2 of periodic_with_pressure_variations (boundary_conditions.py:306-309) */
    const uint8_t *kind_map;   /* nx*ny bytes, reference index order [x][y]; NULL = all fluid */
} lbm_bc_desc;

/* Test hook: compares the kernels' hand-expanded fp64 division (shared reciprocal, lbm_device.cuh) and square root
 * — both the variants that call the library for unusual operands and the branch-free ones of the multi-step kernels,
 * which only flag such operands — with the compiler's IEEE __ddiv_rn / __dsqrt_rn on n generated operand pairs.
 * out[0], out[1] = number of quotients / roots whose bits differ (must be 0), out[2..3] = operands of the first
 * differing quotient, out[4] = operand of the first differing root, out[5], out[6] = how many quotients / roots the
 * branch-free variants answered themselves (not flagged). Guards the bit-exactness of compute_velocity_field
 * (src/lattice_boltzmann_method.py:122-133) and of the norm in equilibrium_distr_func (:185). */
int lbm_selftest_arith(int device, int64_t n, uint64_t seed, uint64_t out[7]);

/* Test hook, host only (no CUDA call): the row plan of the two-steps-per-pass schedule on a lattice with boundary
 * cells. row_has_boundary[x] != 0 marks rows that hold a non-fluid cell. Out: strips[2i], strips[2i+1] = rows [a, b)
 * (b may exceed nx: the strip wraps) advanced by two one-step mask launches through a window; clean[2i], clean[2i+1] =
 * row ranges of the two-step kernel. n_strips = n_clean = -1: the lattice stays on one step per pass (boundary rows
 * within two rows of more than half of all rows, more than 16 ranges, or nx < 16). Arrays need room for 16 ranges. */
int lbm_plan_two_step(int nx, const uint8_t *row_has_boundary, int *n_strips, int *strips, int *n_clean, int *clean);

/* Applies the closures' effect to host arrays (used when a boundary closure is CALLED directly, as the
 * reference's tests/test_boundary_conditions.py does). f_post is updated in place; f_prev may be NULL when no
 * OUTLET rule is present. PBC source flags are ignored here — use lbm_pbc_apply. */
int lbm_bc_apply(int device, int nx, int ny, const lbm_bc_desc *bc, const double *f_pre, double *f_post,
                 const double *f_prev);
/* periodic_with_pressure_variations, x case, in place on f_pre (boundary_conditions.py:337-344). */
int lbm_pbc_apply(int device, int nx, int ny, double rho_in, double rho_out, const double *rho, const double *u,
                  double *f_pre);

/* ---------------------------------------------------------------------------------------------------------
 * Context: a device-resident lattice. nx, ny are the dims of the arrays the reference would hold on this
 * rank (ghost ring included when ghost_x / ghost_y = 1, as experiments.py:618 allocates them).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct lbm_ctx lbm_ctx;

enum {
    LBM_BC_AUTO = 0,      /* flag mask folded into the fused kernel below LBM_EDGE_THRESHOLD cells, edge kernel above */
    LBM_BC_MASK = 1,      /* per-cell flag byte read by the fused kernel */
    LBM_BC_EDGE = 2       /* mask-free periodic kernel over all cells + thin fix-up kernel over the non-fluid cells */
};

/* ghost_x: 0 (periodic wrap in-kernel), 1 (the reference's ghost ring) or 2..4 (slabs of a fluid lattice for the
 * multi-step kernel: g ghost rows per side supply the dependency cone of a g-step pass; needs ghost_y = 0 and no
 * boundary description); ghost_y: 0 or 1. */
int lbm_create(int device, int nx, int ny, int ghost_x, int ghost_y, const lbm_bc_desc *bc /* may be NULL */,
               lbm_ctx **out);
int lbm_destroy(lbm_ctx *ctx);
int lbm_set_bc_mode(lbm_ctx *ctx, int mode);
/* Scheduling switches (A/B measurements, tests):
 *   "fused"          (1) several time steps per pass on bandwidth-bound lattices (temporal blocking): fluid lattices take
 *                    "fused_depth" steps per launch of k_stepNx; on lattices with boundary cells the rows whose two-step
 *                    dependency cone is all fluid take the two-step kernel, the other rows two one-step mask launches
 *                    through a strip window (needs ghost_x = ghost_y = 0, no pressure-periodic rows, and boundary cells
 *                    on at most half of the rows)
 *   "fused_depth"    (3) time steps per pass, 2..4; slabs are limited to their number of ghost rows. A call of n steps is
 *                    n / depth passes, then one pass of the remainder (2 steps: the two-step kernel) or a one-step launch
 *   "deep2"          (0) two-step passes through k_stepNx<2> instead of k_step2x
 *   "graphs"         (1) CUDA-graph replay of 32 captured steps on launch-bound lattices
 *   "pdl"            (1) programmatic dependent launch between the step kernels of launch-bound lattices: the next step's
 *                    blocks are launched and read their parameters / kind bytes while the current step still runs
 *   "generic_kernel" (0) force the one-cell-per-thread step kernel
 *   "tail"           (0) 1 = every call ends with a one-step launch, so that the other buffer holds S_{t-1} (round 1's
 *                    behaviour; A/B and tests). By default a call may END on a multi-step pass: results are then
 *                    materialised by re-running that pass with its last level writing f_post / rho / u instead of
 *                    colliding (fluid rows) or from the strip windows, which keep S_{t-1} of the boundary rows; a changed
 *                    omega redoes the pass's last collision.
 *   "fused_exact"    (0) with "tail" = 1: no tail after all (tests)
 *   "streamed"       (1) lbm_run_host pipelines upload, passes and download over row chunks where it can; 0 = the three calls
 *   "streamed_chunk_rows" (0) rows per chunk of that pipeline (0 = what fits the 256 MB staging buffers; tests)
 *   "wave_seg"       (1) among the long segments pick the one whose block count fills whole waves of resident blocks
 *   "l2_prefetch"    (2) rows ahead of its march whose source segments the multi-step kernel prefetches into L2 with
 *                    cp.async.bulk.prefetch; 0 = off
 *   "fused_seg"      (0) output rows per thread block of the multi-step kernel; 0 = 8..512 by lattice size
 *   "cluster"        (1) lattices that fit the distributed shared memory of one thread-block cluster (16 CTAs, 8 where 16 cannot
 *                    be scheduled; up to ~25 000 cells, no ghost ring) take ALL steps of a call in one launch of k_cluster_steps: the lattice stays in
 *                    shared memory, a step ends with a hardware cluster barrier instead of a kernel boundary. 1 = the first
 *                    four eligible calls are timed alternately on this path and on graph replay and the faster one is kept
 *                    (same bits either way); 2 = always where the lattice fits; 0 = never
 *   "max_queued_calls" (4) lbm_step call k first waits for call k-4 to finish on the device (bounded host run-ahead:
 *                    a driver loop that never reads a result cannot stop its clock with the GPU far behind); 0 = unbounded
 * Environment overrides at lbm_create: LBM_NO_FUSED=1, LBM_NO_GRAPHS=1, LBM_GENERIC_KERNEL=1, LBM_FUSED_SEG=n,
 * LBM_FUSED_DEPTH=d, LBM_DEEP2=1, LBM_NO_CLUSTER=1. */
int lbm_set_option(lbm_ctx *ctx, const char *name, int value);
/* bytes of device memory the context holds */
int64_t lbm_device_bytes(const lbm_ctx *ctx);
/* the cudaStream_t the step kernels are launched on (for CUDA-event timing by the caller) */
void *lbm_stream(lbm_ctx *ctx);

/* Loads the reference's state triple (lattice_boltzmann_method.py:191: f, density, velocity) and performs the
 * first collision f + (feq(rho,u) - f)*omega with the GIVEN moments (:213-215). Synchronous. */
int lbm_upload(lbm_ctx *ctx, const double *f, const double *rho, const double *u, double omega);
/* Device-side initialisation without staging the lattice through the host (initial_values.py:38-123 are all
 * separable): rho(x,y) = rho_x ? rho_x[x] : rho0;  u_x(x,y) = ux_y ? ux_y[y] : ux0;  u_y = uy0;
 * f = feq(rho,u) (lattice_boltzmann_method.py:162-188), then the first collision as in lbm_upload. */
int lbm_init_equilibrium(lbm_ctx *ctx, const double *rho_x, const double *ux_y, double rho0, double ux0,
                         double uy0, double omega);

/* Advances n_steps reference time steps (each: [collide ->] exchange -> stream -> BC -> moments -> collide).
 * omega may differ from the previous call's (experiments.py:171-180 sweeps it): the speculative collision of
 * the last step is then redone from the retained previous buffer. Asynchronous. */
int lbm_step(lbm_ctx *ctx, double omega, int n_steps);
/* Blocks until all queued work of the context is done; reports asynchronous errors. */
int lbm_sync(lbm_ctx *ctx);
/* The whole job from and to host memory: lbm_upload(f, rho, u, omega) + lbm_step(omega, n_steps) + lbm_materialize(f_out,
 * rho_out, u_out) — same results, same final state of the context (the driver loops of src/experiments.py start from host
 * arrays and look at host arrays when they are done: :121-129, :322-327, :756-767). Output pointers may be NULL or alias
 * the inputs. On a fluid lattice without ghost rows the three phases are pipelined over row chunks (time-skewed passes:
 * rows [s_p, X - s_p) of time level p are computable as soon as rows [0, X) have arrived), so that the upload of chunk
 * c+1, the passes over chunk c and the download of the rows chunk c completed overlap on three streams. Slabs with
 * >= steps-per-pass ghost rows and connected neighbours are pipelined too: arrays are the padded local arrays, the
 * interior rows are written; the rows near the slab edges are finished last, pass by pass in lockstep with the
 * neighbours (every rank calls lbm_run_host with the same n_steps; a process-group barrier BEFORE the call, none
 * inside: the upload phase publishes a halo epoch of its own). Other lattices take the three calls one after the
 * other (LBM_ERR_STATE with remote neighbours, where a barrier would be needed in between). Synchronous. */
int lbm_run_host(lbm_ctx *ctx, const double *f, const double *rho, const double *u, double omega, int n_steps,
                 double *f_out, double *rho_out, double *u_out);
/* Number of reference steps taken since the last upload / init. */
int64_t lbm_time(const lbm_ctx *ctx);
/* How many kernels this context has launched so far (bench.py's gpu_launches). */
int64_t lbm_launch_count(const lbm_ctx *ctx);

/* Copies the reference-layout state (f_post, density, velocity of lattice_boltzmann_method.py:225-228) of
 * the current time to host arrays; any pointer may be NULL. Whole arrays, ghost ring included. Synchronous. */
int lbm_materialize(lbm_ctx *ctx, double *f, double *rho, double *u);
/* Same for the sub-rectangle [x0,x1) x [y0,y1) (row-major, packed). */
int lbm_materialize_region(lbm_ctx *ctx, int x0, int x1, int y0, int y1, double *f, double *rho, double *u);

/* Device-side history for drivers that keep a field of EVERY step and look at a few of them after the loop
 * (velocities.append(velocity), src/experiments.py:254, :542): lbm_history_config allocates n_slots slots of
 * (density, velocity) on the device (0 frees them); lbm_history_store parks the fields of the current time in a slot —
 * one asynchronous launch, no copy, no synchronisation; lbm_history_read brings a slot to the host (either pointer may be
 * NULL; synchronous). */
int lbm_history_config(lbm_ctx *ctx, int n_slots);
int lbm_history_store(lbm_ctx *ctx, int slot);
int lbm_history_read(lbm_ctx *ctx, int slot, double *rho, double *u);

/* Probe: records (u_x, u_y) at one cell after every step (experiments.py:703-704) into a ring in host-mapped
 * memory — the cell's thread stores the sample, a system-scope fence and the step number straight to the host as
 * the step completes. lbm_probe_read copies the samples of steps [t0, t0+n) — at most `capacity` behind
 * lbm_time() — and waits only until step t0+n-1 has reported: steps queued behind it keep running (choose
 * capacity >= the number of steps queued ahead of the reader). */
int lbm_probe_config(lbm_ctx *ctx, int x, int y, int capacity);
int lbm_probe_read(lbm_ctx *ctx, int64_t t0, int n, double *uxuy);
/* Whole-field extrema of the current state (experiments.py:181-193): out = {min rho, max rho, min u, max u}
 * over the cells [x0,x1) x [y0,y1); u extrema are over both components, as np.amin(velocity). Synchronous. */
int lbm_minmax(lbm_ctx *ctx, int x0, int x1, int y0, int y1, double out[4]);

/* ---------------------------------------------------------------------------------------------------------
 * Halo exchange (replaces communication(), src/parallelization_utils.py:6-52). Ghost cells of a neighbour
 * are written DIRECTLY by the kernel that computes the edge cells (peer stores over NVLink through a CUDA-IPC
 * mapping); there is no separate copy or pack step. Ordering: the last block of every ghost-storing kernel of
 * step n publishes "done n" into each remote neighbour's flag word; a ghost-touching kernel of step n+1 first
 * waits until all its remote neighbours are done with n. All ranks must therefore call lbm_step in lockstep
 * (same n_steps), as MPI ranks call Sendrecv in lockstep; ranks may drift by at most one step. A neighbour that
 * never arrives makes the wait time out (LBM_HALO_TIMEOUT_S, default 30 s): lbm_sync then returns
 * LBM_ERR_TIMEOUT. Loads (lbm_upload / lbm_init_equilibrium) also store ghosts: bracket them with a
 * process-group barrier on both sides.
 * Neighbour slot index = (dx+1)*3 + (dy+1), dx,dy in {-1,0,1}, (0,0) unused.
 * ------------------------------------------------------------------------------------------------------- */
#define LBM_IPC_HANDLE_BYTES 64
typedef struct {
    uint8_t mem_handle[LBM_IPC_HANDLE_BYTES];  /* cudaIpcMemHandle_t of the context's arena */
    int32_t device;
    int32_t nx, ny, pitch;
    int64_t pid;                               /* same pid => same process: use the pointer directly */
    uint64_t arena_ptr;                        /* only meaningful within that process */
    int64_t arena_bytes;
} lbm_halo_export;

int lbm_halo_export_handle(lbm_ctx *ctx, lbm_halo_export *out);
/* slot's neighbour is the context described by `peer` (may be this context itself: self-periodic wrap). */
int lbm_halo_connect(lbm_ctx *ctx, int slot, const lbm_halo_export *peer);
/* Call once after every rank has connected all its neighbours (and after a process-group barrier). */
int lbm_halo_finalize(lbm_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
