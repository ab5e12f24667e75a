// lbm_b200.cu — B200 (sm_100a) implementation of include/lbm_b200.h: the host side (contexts, scheduling, C-ABI).
// The kernels are in lbm_kernels.cuh, the per-cell arithmetic in lbm_device.cuh.
//
// State on the device is the POST-collision population set S[i][x][y] (SoA fp64, rows padded to 128 B, two
// buffers A/B). One fused kernel maps S_t -> S_{t+1} (or S_{t+D}) per launch:
//     pull-stream (lattice_boltzmann_method.py:153-157) -> boundary rules (boundary_conditions.py) ->
//     moments (:93-137) -> equilibrium (:162-188) -> BGK collide (:215) -> store [+ ghost stores to neighbours]
// reading every population once and writing it once per launch (144 B per cell update, 144 / D with D steps per pass).
// See DESIGN.md.
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <unistd.h>

#include "../../include/lbm_b200.h"
#include "lbm_device.cuh"

using namespace lbm;

// -------------------------------------------------------------------------------------------------------
// errors
// -------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(LBM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char *lbm_last_error(void) { return g_err.c_str(); }
extern "C" const char *lbm_version(void) { return "lbm_b200 0.1 (sm_100a, fp64, pull/post-collision state)"; }
extern "C" int lbm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

#include "lbm_kernels.cuh"

// -------------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------------
struct Peer {
    bool connected = false, remote = false;
    void *mapped = nullptr;     // cudaIpcOpenMemHandle result (remote only)
    char *arena = nullptr;      // base of the neighbour's arena in this process' address space
    int nx = 0, ny = 0, pitch = 0;
};

struct lbm_ctx {
    int device = 0;
    int NX = 0, NY = 0, pitch = 0, gx = 0, gy = 0;
    long long plane = 0;
    // arena layout (one allocation so that one IPC handle exports everything a neighbour needs)
    char *arena = nullptr;
    size_t arena_bytes = 0;
    size_t off_S[2] = {0, 0}, off_flags = 0;
    double *S[2] = {nullptr, nullptr};
    unsigned *flags_in = nullptr;       // [9] inside the arena
    // other device memory
    uint8_t *kind_map = nullptr;
    lbm_kind *kinds = nullptr;
    double *ktab = nullptr, *ctab = nullptr;
    double *outbuf[4] = {nullptr, nullptr, nullptr, nullptr};   // outlet side buffers of S[0], S[1] and of the strips' windows (2: the last one, 3: the first of three-step passes)
    double *snap_row = nullptr, *snap_col = nullptr;
    // Two steps per pass on lattices WITH boundary cells (plan_strips): rows whose two-step dependency cone holds
    // fluid cells only go through k_step2x; the few others are advanced by two one-step mask launches through a
    // window that holds the intermediate state S_{t+1} of the strip (+ one row each side).
    struct Strip {
        int a, b;              // output rows [a, b), unwrapped: 0 <= a < NX, b may exceed NX (periodic wrap)
        double *buf;           // [9][b - a + 2][pitch]: rows a-1 .. b of the LAST intermediate state (S_{t+1} of a two-step, S_{t+2} of a three-step pass)
        double *buf1;          // [9][b - a + 4][pitch]: rows a-2 .. b+1 of S_{t+1} of a three-step pass (null when planned for two steps)
    };
    int bc_depth = 2;          // steps per pass the strips were planned for (reach of a boundary row: +- bc_depth rows)
    std::vector<Strip> strips;
    std::vector<std::pair<int, int>> clean;   // row ranges [first, last) of k_step2x
    bool fused_bc = false;
    int2 *cells = nullptr;
    int n_cells = 0;
    bool has_bc = false;
    double rho_in = 0, rho_out = 0;
    int bc_mode = LBM_BC_AUTO;
    bool force_generic = false;   // LBM_GENERIC_KERNEL=1: always use the one-cell-per-thread kernel (A/B measurements)
    unsigned *done_counter = nullptr, *err_flag = nullptr;
    unsigned *err_host = nullptr;   // cudaHostAlloc (mapped): set by a kernel whose halo wait timed out
    long long timeout_cycles = 30LL * 2000000000LL;   // ~30 s at 2 GHz; LBM_HALO_TIMEOUT_S overrides
    // staging for upload / materialize (reference layout, a chunk of rows)
    double *stage_f = nullptr, *stage_rho = nullptr, *stage_u = nullptr;
    long long stage_cells = 0;
    long long *mm_acc = nullptr;
    // probe ring
    int px = -1, py = -1, probe_cap = 0;
    double *probe = nullptr;         // cudaHostAlloc (mapped): written by the probe cell's thread, read by the host
    long long *progress = nullptr;   // cudaHostAlloc (mapped)
    long long *tcount = nullptr;   // device [4]: time of the state in S[0], S[1], and in the strip windows (2: last, 3: first of three)
    // CUDA graphs of kGraphSteps steps for launch-bound lattices, keyed by (omega, parity, probe, bc mode)
    struct GraphEntry {
        double omega;
        int parity, mode;
        bool snap_last;      // the last captured step keeps the ghost snapshot (a call that ENDS on this replay)
        const void *probe;
        cudaGraphExec_t exec;
        long long launches;
    };
    std::vector<GraphEntry> graphs;
    // The ghost-ring snapshot (snapshot_ghosts) serves the materialisation after a call's LAST step only: every other
    // step of a call skips it (a load round trip + stores on the edge threads of a launch-bound step).
    bool skip_snap = false;
    bool eager_progress = false;   // lbm_step(.., 1): the probe sample's time word is published by the step itself
    bool use_graphs = true;
    bool pdl = true;              // programmatic dependent launch between the step kernels of launch-bound lattices (option "pdl")
    bool use_fused = true;        // LBM_NO_FUSED=1: one step per pass only (A/B measurements)
    int streamed_chunk_rows = 0;  // lbm_run_host: rows per chunk (0 = what fits the 256 MB staging buffer; option "streamed_chunk_rows")
    bool use_streamed = true;     // lbm_run_host pipelines upload / passes / download over row chunks (option "streamed")
    bool use_cluster = true;      // lattices that fit a thread-block cluster's shared memory: many steps per launch (option "cluster")
    int cluster_size = 0;         // 0 = not decided yet, -1 = not possible on this lattice / device, else CTAs per cluster
    int cluster_rows = 0, cluster_threads = 0, cluster_m = 0;
    size_t cluster_smem = 0;
    // cluster kernel or graph replay? Both give the same bits; which one is faster depends on how the rows divide over
    // the cluster's CTAs and on the boundary cells (profiles/r02_cluster_vs_graph.txt), so the first eligible calls are
    // TIMED (CUDA events around real steps, read back without blocking) alternately on both paths and the faster one
    // is kept: tune_pick = 0 undecided, 1 cluster, 2 graphs.
    int tune_pick = 0, tune_calls = 0;
    struct TuneSample {
        cudaEvent_t a = nullptr, b = nullptr;
        int steps = 0, path = 0;
        bool pending = false;
    } tune[4];
    int l2_prefetch = 2;          // rows ahead whose source segments k_step2x prefetches into L2 (option "l2_prefetch")
    int fused_seg = 0;            // output rows per block of the two-step kernel; 0 = pick_seg (LBM_FUSED_SEG / option "fused_seg")
    int n_sm = 148;               // multiprocessors of the device (wave-aware segment length)
    bool wave_seg = true;         // option "wave_seg": pick the segment length whose block count fills whole waves
    int fused_depth = 3;          // time steps per pass of the multi-step kernel, 2..4 (LBM_FUSED_DEPTH / option "fused_depth")
    bool deep2 = false;           // depth-2 passes through k_stepNx<2> (18-slot ring) instead of k_step2x (option "deep2")
    bool fused_exact = false;     // tests: an even lbm_step(n) is exactly n/2 two-step passes (no one-step tail)
    bool force_tail = false;      // end every call with a one-step launch even on fluid lattices (option "tail": ranks of one
                                  // decomposition must take the same launch sequence, and those with boundary cells need the tail)
    int last_depth = 0;           // S[cur^1] holds S_{t-last_depth}: depth of the pass that produced S_t (0 right after a load)
    // state
    int cur = 0;              // S[cur] = S_t
    bool loaded = false;
    long long t = 0;          // reference steps since upload
    double omega = 0;
    long long launches = 0;
    unsigned halo_epoch = 0;  // value published after the kernel that produced S_t
    // halo
    Peer peer[9];
    bool halo_ready = false, any_remote = false;
    cudaStream_t stream = nullptr, stream_edge = nullptr;
    cudaEvent_t ev_main = nullptr, ev_edge = nullptr;
    // device-side history of (rho, u) fields (lbm_history_*): n_hist slots of 3 * NX * NY doubles
    double *hist = nullptr;
    int n_hist = 0;
    // lbm_run_host: double-buffered staging (set 0 of the inputs is stage_f / stage_rho / stage_u), copy streams, events
    struct Streamed {
        double *in_f[2] = {}, *in_rho[2] = {}, *in_u[2] = {}, *out_f[2] = {}, *out_rho[2] = {}, *out_u[2] = {};
        cudaStream_t h2d = nullptr, d2h = nullptr;
        cudaEvent_t in_ready[2] = {}, in_free[2] = {}, out_ready[2] = {}, out_free[2] = {};
        long long *tclock = nullptr;   // device: time after pass p (probe clock of the skewed schedule)
        int tclock_cap = 0;
    } *sr = nullptr;
    // Bounded run-ahead of the host (option "max_queued_calls", default 4): lbm_step call k first waits until call
    // k - max_queued_calls has finished on the device. A driver that never looks at a result inside its loop (the
    // reference's scaling_test stops its clock right after the loop, src/experiments.py:763-769) then cannot run more
    // than a few batches ahead of the GPU, so its wall clock covers the work it claims.
    static const int kCallRing = 16;
    cudaEvent_t ev_call[kCallRing] = {};
    long long n_calls = 0;
    int max_queued_calls = 4;
};

static void drop_graphs(lbm_ctx *c);
static const int kFusedThreads = 128;   // two-steps-per-pass kernel: threads per block
#ifndef LBM_DEEP_T
#define LBM_DEEP_T 128
#endif
static const int kDeepThreads = LBM_DEEP_T;   // k_stepNx: threads per block (NY >= 2 * max of the two is required)
static const long long kEdgeThreshold = 1 << 20;   // cells; LBM_BC_AUTO switches to the edge kernel above this
static const int kMaxDepth = 4;         // deepest multi-step pass (k_stepNx)

typedef void (*deep_fn)(const StepParams);
template <int D>
static deep_fn deep_kernel_d(bool halo, bool probe, bool final)
{
    constexpr int T = kDeepThreads;
    if (final) return k_stepNx<T, D, false, false, true>;
    if (halo) return probe ? k_stepNx<T, D, true, true, false> : k_stepNx<T, D, true, false, false>;
    return probe ? k_stepNx<T, D, false, true, false> : k_stepNx<T, D, false, false, false>;
}
static deep_fn deep_kernel(int depth, bool halo, bool probe, bool final)
{
    return depth == 2 ? deep_kernel_d<2>(halo, probe, final) : (depth == 3 ? deep_kernel_d<3>(halo, probe, final) : deep_kernel_d<4>(halo, probe, final));
}
static int deep_smem(int depth) { return (depth - 1) * 18 * 2 * kDeepThreads * (int)sizeof(double); }
static int deep_width(int depth) { return 2 * kDeepThreads - 4 * (depth - 1); }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int set_device(int device)
{
    CK(cudaSetDevice(device));
    return LBM_OK;
}

// ---- stateless ops -------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
    template <class T>
    T *as()
    {
        return (T *)p;
    }
};

extern "C" int lbm_equilibrium(int device, int64_t n, const double *rho, const double *u, double *f_out)
{
    if (n < 0 || !rho || !u || !f_out) return fail(LBM_ERR_ARG, "lbm_equilibrium: null pointer or negative size");
    if (n == 0) return LBM_OK;
    if (int rc = set_device(device)) return rc;
    DevBuf dr, du, df;
    CK(dr.alloc(n * 8));
    CK(du.alloc(n * 16));
    CK(df.alloc(n * 72));
    CK(cudaMemcpy(dr.p, rho, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(du.p, u, n * 16, cudaMemcpyHostToDevice));
    k_equilibrium<<<(unsigned)((n + 255) / 256), 256>>>(n, dr.as<double>(), du.as<double>(), df.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpy(f_out, df.p, n * 72, cudaMemcpyDeviceToHost));
    return LBM_OK;
}

static int moments_op(int device, int64_t n, const double *rho_in, const double *f, double *rho_out, double *u_out)
{
    if (n == 0) return LBM_OK;
    if (int rc = set_device(device)) return rc;
    DevBuf df, dri, dro, du;
    CK(df.alloc(n * 72));
    CK(cudaMemcpy(df.p, f, n * 72, cudaMemcpyHostToDevice));
    if (rho_in) {
        CK(dri.alloc(n * 8));
        CK(cudaMemcpy(dri.p, rho_in, n * 8, cudaMemcpyHostToDevice));
    }
    if (rho_out) CK(dro.alloc(n * 8));
    if (u_out) CK(du.alloc(n * 16));
    k_moments<<<(unsigned)((n + 255) / 256), 256>>>(n, df.as<double>(), rho_in ? dri.as<double>() : nullptr,
                                                    rho_out ? dro.as<double>() : nullptr, u_out ? du.as<double>() : nullptr);
    CK(cudaGetLastError());
    if (rho_out) CK(cudaMemcpy(rho_out, dro.p, n * 8, cudaMemcpyDeviceToHost));
    if (u_out) CK(cudaMemcpy(u_out, du.p, n * 16, cudaMemcpyDeviceToHost));
    return LBM_OK;
}

extern "C" int lbm_density(int device, int64_t n, const double *f, double *rho_out)
{
    if (n < 0 || !f || !rho_out) return fail(LBM_ERR_ARG, "lbm_density: null pointer or negative size");
    return moments_op(device, n, nullptr, f, rho_out, nullptr);
}

extern "C" int lbm_velocity(int device, int64_t n, const double *rho, const double *f, double *u_out)
{
    if (n < 0 || !rho || !f || !u_out) return fail(LBM_ERR_ARG, "lbm_velocity: null pointer or negative size");
    return moments_op(device, n, rho, f, nullptr, u_out);
}

extern "C" int lbm_streaming(int device, int nx, int ny, const double *f, double *f_out)
{
    if (nx <= 0 || ny <= 0 || !f || !f_out) return fail(LBM_ERR_ARG, "lbm_streaming: bad shape or null pointer");
    if (int rc = set_device(device)) return rc;
    const long long n = (long long)nx * ny;
    DevBuf a, b;
    CK(a.alloc(n * 72));
    CK(b.alloc(n * 72));
    CK(cudaMemcpy(a.p, f, n * 72, cudaMemcpyHostToDevice));
    k_streaming_aos<<<(unsigned)((n + 255) / 256), 256>>>(nx, ny, a.as<double>(), b.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpy(f_out, b.p, n * 72, cudaMemcpyDeviceToHost));
    return LBM_OK;
}

extern "C" int lbm_selftest_arith(int device, int64_t n, uint64_t seed, uint64_t out[7])
{
    if (n < 0 || !out) return fail(LBM_ERR_ARG, "lbm_selftest_arith: bad argument");
    if (int rc = set_device(device)) return rc;
    DevBuf d;
    CK(d.alloc(7 * 8));
    CK(cudaMemset(d.p, 0, 7 * 8));
    if (n) k_selftest_arith<<<(unsigned)((n + 255) / 256), 256>>>(n, seed, d.as<unsigned long long>());
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, d.p, 7 * 8, cudaMemcpyDeviceToHost));
    return LBM_OK;
}

static int check_bc(const lbm_bc_desc *bc)
{
    if (bc->n_kinds < 1 || bc->n_kinds > 256 || !bc->kinds) return fail(LBM_ERR_ARG, "bc: n_kinds must be 1..256");
    if (bc->n_k_rows < 1 || bc->n_k_rows > 32 || !bc->k_table) return fail(LBM_ERR_ARG, "bc: n_k_rows must be 1..32");
    if (bc->n_c_rows < 0 || bc->n_c_rows > 32) return fail(LBM_ERR_ARG, "bc: n_c_rows must be 0..32");
    for (int i = 0; i < 9; i++) {
        if (bc->kinds[0].rule[i] != 0) return fail(LBM_ERR_ARG, "bc: kind 0 must be the all-PULL fluid cell");
        if (bc->k_table[i] != 0.0) return fail(LBM_ERR_ARG, "bc: row 0 of k_table must be zero");
    }
    if (bc->kinds[0].flags || bc->kinds[0].skip_store) return fail(LBM_ERR_ARG, "bc: kind 0 must carry no flags");
    for (int k = 0; k < bc->n_kinds; k++)
        for (int i = 0; i < 9; i++) {
            const int type = bc->kinds[k].rule[i] & 7, row = bc->kinds[k].rule[i] >> 3;
            if (type > LBM_RULE_OUTLET) return fail(LBM_ERR_ARG, "bc: unknown rule type %d", type);
            if (type == LBM_RULE_BOUNCE && row >= bc->n_k_rows) return fail(LBM_ERR_ARG, "bc: k_table row out of range");
            if (type == LBM_RULE_CONST && row >= bc->n_c_rows) return fail(LBM_ERR_ARG, "bc: c_table row out of range");
            if (type == LBM_RULE_OUTLET && !(i == 3 || i == 6 || i == 7))
                return fail(LBM_ERR_ARG, "bc: OUTLET rule only exists for populations 3, 6, 7");
        }
    return LBM_OK;
}

extern "C" int lbm_bc_apply(int device, int nx, int ny, const lbm_bc_desc *bc, const double *f_pre, double *f_post,
                            const double *f_prev)
{
    if (nx <= 0 || ny <= 0 || !bc || !f_pre || !f_post || !bc->kind_map) return fail(LBM_ERR_ARG, "lbm_bc_apply: bad argument");
    if (int rc = check_bc(bc)) return rc;
    if (int rc = set_device(device)) return rc;
    const long long n = (long long)nx * ny;
    DevBuf dm, dk, dkt, dct, a, b, c;
    CK(dm.alloc(n));
    CK(dk.alloc(bc->n_kinds * sizeof(lbm_kind)));
    CK(dkt.alloc(bc->n_k_rows * 72));
    CK(dct.alloc(std::max(bc->n_c_rows, 1) * 72));
    CK(a.alloc(n * 72));
    CK(b.alloc(n * 72));
    CK(cudaMemcpy(dm.p, bc->kind_map, n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dk.p, bc->kinds, bc->n_kinds * sizeof(lbm_kind), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dkt.p, bc->k_table, bc->n_k_rows * 72, cudaMemcpyHostToDevice));
    if (bc->n_c_rows) CK(cudaMemcpy(dct.p, bc->c_table, bc->n_c_rows * 72, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(a.p, f_pre, n * 72, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b.p, f_post, n * 72, cudaMemcpyHostToDevice));
    if (f_prev) {
        CK(c.alloc(n * 72));
        CK(cudaMemcpy(c.p, f_prev, n * 72, cudaMemcpyHostToDevice));
    } else {
        for (int k = 0; k < bc->n_kinds; k++)
            for (int i = 0; i < 9; i++)
                if ((bc->kinds[k].rule[i] & 7) == LBM_RULE_OUTLET) return fail(LBM_ERR_ARG, "lbm_bc_apply: OUTLET rule needs f_prev");
    }
    k_bc_apply_aos<<<(unsigned)((n + 255) / 256), 256>>>(nx, ny, dm.as<uint8_t>(), dk.as<lbm_kind>(), dkt.as<double>(),
                                                         dct.as<double>(), a.as<double>(), b.as<double>(), c.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpy(f_post, b.p, n * 72, cudaMemcpyDeviceToHost));
    return LBM_OK;
}

extern "C" int lbm_pbc_apply(int device, int nx, int ny, double rho_in, double rho_out, const double *rho, const double *u,
                             double *f_pre)
{
    if (nx < 4 || ny <= 0 || !rho || !u || !f_pre) return fail(LBM_ERR_ARG, "lbm_pbc_apply: needs nx >= 4 and non-null arrays");
    if (int rc = set_device(device)) return rc;
    const long long n = (long long)nx * ny;
    DevBuf dr, du, df;
    CK(dr.alloc(n * 8));
    CK(du.alloc(n * 16));
    CK(df.alloc(n * 72));
    CK(cudaMemcpy(dr.p, rho, n * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(du.p, u, n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(df.p, f_pre, n * 72, cudaMemcpyHostToDevice));
    k_pbc_apply_aos<<<(ny + 255) / 256, 256>>>(nx, ny, rho_in, rho_out, dr.as<double>(), du.as<double>(), df.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpy(f_pre, df.p, n * 72, cudaMemcpyDeviceToHost));
    return LBM_OK;
}

// ---- CUDA-IPC mappings ---------------------------------------------------------------------------------------
// cudaIpcOpenMemHandle may be called once per process and handle: mappings are shared between the neighbour slots (and
// contexts) that name the same peer arena, counted, and closed when the last user lets go — a later connect with the
// same handle bytes (re-attach after release_lattices / eviction) opens a fresh mapping instead of finding a dead one.
struct IpcMapping {
    std::string key;
    void *ptr;
    int refs;
};
static std::mutex g_ipc_mutex;
static std::vector<IpcMapping> g_ipc;

static int ipc_acquire(const uint8_t *handle_bytes, void **out)
{
    std::lock_guard<std::mutex> lock(g_ipc_mutex);
    const std::string key((const char *)handle_bytes, LBM_IPC_HANDLE_BYTES);
    for (auto &m : g_ipc)
        if (m.key == key) {
            m.refs++;
            *out = m.ptr;
            return LBM_OK;
        }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_bytes, sizeof h);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g_ipc.push_back({key, p, 1});
    *out = p;
    return LBM_OK;
}

static void ipc_release(void *ptr)
{
    std::lock_guard<std::mutex> lock(g_ipc_mutex);
    for (size_t i = 0; i < g_ipc.size(); i++)
        if (g_ipc[i].ptr == ptr) {
            if (--g_ipc[i].refs == 0) {
                cudaIpcCloseMemHandle(ptr);
                g_ipc.erase(g_ipc.begin() + i);
            }
            return;
        }
}

// ---- context -------------------------------------------------------------------------------------------
extern "C" int lbm_destroy(lbm_ctx *c)
{
    if (!c) return LBM_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream_edge) cudaStreamSynchronize(c->stream_edge);
    drop_graphs(c);
    for (int s = 0; s < 9; s++)
        if (c->peer[s].mapped) ipc_release(c->peer[s].mapped);   // every slot holds its own reference
    for (auto &s : c->strips) {
        if (s.buf) cudaFree(s.buf);
        if (s.buf1) cudaFree(s.buf1);
    }
    void *bufs[] = {c->arena,  c->kind_map, c->kinds,    c->ktab,   c->ctab,  c->outbuf[0],   c->outbuf[1], c->outbuf[2], c->outbuf[3], c->cells, c->snap_row, c->snap_col,
                    c->done_counter, c->err_flag, c->stage_f, c->stage_rho, c->stage_u, c->mm_acc, c->tcount};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (c->hist) cudaFree(c->hist);
    if (c->probe) cudaFreeHost(c->probe);
    if (c->progress) cudaFreeHost(c->progress);
    if (c->err_host) cudaFreeHost(c->err_host);
    if (c->sr) {
        for (int k = 0; k < 2; k++) {
            if (k) { cudaFree(c->sr->in_f[k]); cudaFree(c->sr->in_rho[k]); cudaFree(c->sr->in_u[k]); }
            cudaFree(c->sr->out_f[k]); cudaFree(c->sr->out_rho[k]); cudaFree(c->sr->out_u[k]);
            for (cudaEvent_t e : {c->sr->in_ready[k], c->sr->in_free[k], c->sr->out_ready[k], c->sr->out_free[k]})
                if (e) cudaEventDestroy(e);
        }
        if (c->sr->tclock) cudaFree(c->sr->tclock);
        if (c->sr->h2d) cudaStreamDestroy(c->sr->h2d);
        if (c->sr->d2h) cudaStreamDestroy(c->sr->d2h);
        delete c->sr;
    }
    for (cudaEvent_t e : c->ev_call)
        if (e) cudaEventDestroy(e);
    for (auto &t : c->tune) {
        if (t.a) cudaEventDestroy(t.a);
        if (t.b) cudaEventDestroy(t.b);
    }
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_edge) cudaEventDestroy(c->ev_edge);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream_edge) cudaStreamDestroy(c->stream_edge);
    delete c;
    return LBM_OK;
}

// Which rows may take two time steps in one pass on a lattice with boundary cells? Row r of S_{t+2} is CLEAN when
// rows r-2 .. r+2 (periodic) hold fluid cells only: then the intermediate rows r-1 .. r+1 and row r itself are plain
// pulls, which is all k_step2x knows. Maximal runs of the other rows are STRIPS; a strip [a, b) is advanced by the
// one-step mask kernel twice, S_t rows [a-2, b+2) -> window rows [a-1, b+1) of S_{t+1} -> S_{t+2} rows [a, b).
// For the von Karman rule set (inlet row, plate rows, outlet rows) 13 of the NX rows are strip rows.
// (pure host logic; lbm_plan_two_step exposes it to the CPU tests)
// Slabs (gx >= 2 ghost rows per side, no periodic wrap inside the array): the rows this rank computes are [gx, NX-gx);
// a row is a strip row when a non-fluid cell — on a ghost row too — lies within two rows, and ALWAYS within gx rows of
// the slab edges: those are the rows that read ghost rows and whose results are stored into the neighbours' ghost rows,
// which on a lattice with boundary cells is done by the one-step mask kernel (k_step2x knows neither rules nor halos
// ... the fluid edge launch of two_steps() does, but not next to boundary cells; one code path is enough here).
static bool plan_rows_slab(int NX, int gx, const std::vector<char> &dirty, std::vector<std::pair<int, int>> &strips,
                           std::vector<std::pair<int, int>> &clean, int reach = 2)
{
    strips.clear();
    clean.clear();
    const int lo = gx, hi = NX - gx;
    if (hi - lo < 8 * gx) return false;
    std::vector<char> strip_row(NX, 0);
    int n_strip_rows = 0;
    for (int x = lo; x < hi; x++) {
        for (int d = -reach; d <= reach; d++) strip_row[x] |= dirty[x + d];
        if (x < lo + gx || x >= hi - gx) strip_row[x] = 1;
        n_strip_rows += strip_row[x];
    }
    if (2 * n_strip_rows > hi - lo) return false;
    for (int x = lo; x < hi;) {
        int e = x;
        while (e < hi && strip_row[e] == strip_row[x]) e++;
        (strip_row[x] ? strips : clean).push_back({x, e});
        x = e;
    }
    return strips.size() <= 16 && clean.size() <= 16;
}

static bool plan_rows(int NX, const std::vector<char> &dirty, std::vector<std::pair<int, int>> &strips,
                      std::vector<std::pair<int, int>> &clean, int reach = 2)
{
    strips.clear();
    clean.clear();
    if (NX < 16) return false;
    std::vector<char> strip_row(NX, 0);
    int n_strip_rows = 0;
    for (int x = 0; x < NX; x++) {
        for (int d = -reach; d <= reach; d++) strip_row[x] |= dirty[((x + d) % NX + NX) % NX];
        n_strip_rows += strip_row[x];
    }
    if (2 * n_strip_rows > NX) return false;   // mostly boundary rows (walls along x, ...): one step per pass
    for (int a = 0; a < NX; a++) {
        if (!strip_row[a] || strip_row[(a + NX - 1) % NX]) continue;   // a strip begins after a clean row
        int b = a;
        while (strip_row[b % NX]) b++;
        strips.push_back({a, b});   // b may exceed NX: the strip wraps
    }
    for (int x = 0; x < NX;) {
        if (strip_row[x]) {
            x++;
            continue;
        }
        int e = x;
        while (e < NX && !strip_row[e]) e++;
        clean.push_back({x, e});
        x = e;
    }
    return strips.size() <= 16 && clean.size() <= 16;
}

extern "C" int lbm_plan_two_step(int nx, const uint8_t *row_has_boundary, int *n_strips, int *strips, int *n_clean, int *clean)
{
    if (nx < 1 || !row_has_boundary || !n_strips || !strips || !n_clean || !clean) return fail(LBM_ERR_ARG, "lbm_plan_two_step: bad argument");
    std::vector<char> dirty(row_has_boundary, row_has_boundary + nx);
    std::vector<std::pair<int, int>> st, cl;
    if (!plan_rows(nx, dirty, st, cl)) {
        *n_strips = *n_clean = -1;
        return LBM_OK;
    }
    *n_strips = (int)st.size();
    *n_clean = (int)cl.size();
    for (size_t i = 0; i < st.size(); i++) {
        strips[2 * i] = st[i].first;
        strips[2 * i + 1] = st[i].second;
    }
    for (size_t i = 0; i < cl.size(); i++) {
        clean[2 * i] = cl[i].first;
        clean[2 * i + 1] = cl[i].second;
    }
    return LBM_OK;
}

static int plan_strips(lbm_ctx *c, const std::vector<char> &dirty)
{
    // planned for the deepest pass this lattice may take (a shallower pass can use the same, wider strips)
    const int reach = std::max(2, std::min(3, c->gx ? std::min(c->gx, c->fused_depth) : c->fused_depth));
    std::vector<std::pair<int, int>> rows, clean;
    if (!(c->gx ? plan_rows_slab(c->NX, c->gx, dirty, rows, clean, reach) : plan_rows(c->NX, dirty, rows, clean, reach))) return LBM_OK;
    for (const auto &r : rows) {
        lbm_ctx::Strip s = {r.first, r.second, nullptr, nullptr};
        const size_t bytes = (size_t)9 * (s.b - s.a + 2) * c->pitch * 8, bytes1 = (size_t)9 * (s.b - s.a + 4) * c->pitch * 8;
        if (cudaMalloc(&s.buf, bytes) != cudaSuccess || (reach == 3 && cudaMalloc(&s.buf1, bytes1) != cudaSuccess)) {
            cudaGetLastError();
            if (s.buf) cudaFree(s.buf);
            return fail(LBM_ERR_NOMEM, "cannot allocate %.1f MB for a boundary strip window", (bytes + bytes1) / 1e6);
        }
        c->strips.push_back(s);   // owned by the context from here on (lbm_destroy frees it)
        CK(cudaMemsetAsync(s.buf, 0, bytes, c->stream));
        if (s.buf1) CK(cudaMemsetAsync(s.buf1, 0, bytes1, c->stream));
    }
    c->clean = clean;
    c->fused_bc = true;
    c->bc_depth = reach;
    return LBM_OK;
}

static int ctx_build(lbm_ctx *c, const lbm_bc_desc *bc)
{
    CK(cudaSetDevice(c->device));
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, c->device));
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&c->stream_edge, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_edge, cudaEventDisableTiming));
    for (cudaEvent_t &e : c->ev_call) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    {
        const int smem = 4 * 6 * 2 * kFusedThreads * (int)sizeof(double);
        CK(cudaFuncSetAttribute(k_step2x<kFusedThreads, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k_step2x<kFusedThreads, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k_step2x<kFusedThreads, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(k_step2x<kFusedThreads, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int d = 2; d <= kMaxDepth; d++)
            for (int v = 0; v < 5; v++)
                CK(cudaFuncSetAttribute((const void *)deep_kernel(d, v & 1, (v >> 1) & 1, v == 4), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        deep_smem(d)));
    }

    const size_t sbytes = (size_t)9 * c->plane * 8;
    c->off_S[0] = 0;
    c->off_S[1] = align_up(sbytes, 256);
    c->off_flags = c->off_S[1] + align_up(sbytes, 256);
    c->arena_bytes = c->off_flags + 256;
    if (cudaMalloc(&c->arena, c->arena_bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(LBM_ERR_NOMEM, "cannot allocate %.2f GB for a %dx%d lattice (two fp64 SoA buffers)", c->arena_bytes / 1e9,
                    c->NX, c->NY);
    }
    c->S[0] = (double *)(c->arena + c->off_S[0]);
    c->S[1] = (double *)(c->arena + c->off_S[1]);
    c->flags_in = (unsigned *)(c->arena + c->off_flags);
    CK(cudaMemsetAsync(c->arena + c->off_flags, 0, 256, c->stream));
    // padding columns are never read, but keep the buffers defined
    CK(cudaMemsetAsync(c->S[0], 0, sbytes, c->stream));
    CK(cudaMemsetAsync(c->S[1], 0, sbytes, c->stream));
    CK(cudaMalloc(&c->done_counter, 8));
    CK(cudaMalloc(&c->err_flag, 4));
    CK(cudaMemsetAsync(c->done_counter, 0, 8, c->stream));
    CK(cudaMemsetAsync(c->err_flag, 0, 4, c->stream));
    CK(cudaMalloc(&c->mm_acc, 32));
    CK(cudaMalloc(&c->tcount, 32));
    CK(cudaMemsetAsync(c->tcount, 0, 32, c->stream));
    CK(cudaHostAlloc((void **)&c->progress, 64, cudaHostAllocMapped));
    *c->progress = 0;
    CK(cudaHostAlloc((void **)&c->err_host, 64, cudaHostAllocMapped));
    *c->err_host = 0;
    for (int b = 0; b < 4; b++) {
        CK(cudaMalloc(&c->outbuf[b], (size_t)3 * c->pitch * 8));
        CK(cudaMemsetAsync(c->outbuf[b], 0, (size_t)3 * c->pitch * 8, c->stream));
    }

    if (c->gx || c->gy) {
        CK(cudaMalloc(&c->snap_row, (size_t)2 * 9 * c->pitch * 8));
        CK(cudaMemsetAsync(c->snap_row, 0, (size_t)2 * 9 * c->pitch * 8, c->stream));
        CK(cudaMalloc(&c->snap_col, (size_t)2 * 9 * c->NX * 8));
        CK(cudaMemsetAsync(c->snap_col, 0, (size_t)2 * 9 * c->NX * 8, c->stream));
    }

    // staging: a chunk of rows in reference layout (96 B per cell), at most ~256 MB
    long long rows = std::max<long long>(1, std::min<long long>(c->NX, (256LL << 20) / (96LL * c->NY)));
    c->stage_cells = rows * c->NY;
    CK(cudaMalloc(&c->stage_f, c->stage_cells * 72));
    CK(cudaMalloc(&c->stage_rho, c->stage_cells * 8));
    CK(cudaMalloc(&c->stage_u, c->stage_cells * 16));

    if (bc && bc->kind_map) {
        if (int rc = check_bc(bc)) return rc;
        c->has_bc = true;
        c->rho_in = bc->pbc_rho_in;
        c->rho_out = bc->pbc_rho_out;
        std::vector<uint8_t> km((size_t)c->NX * c->pitch, 0);
        std::vector<int2> cells;            // non-fluid cells this rank computes (fix-up list of the edge kernel)
        std::vector<char> dirty(c->NX, 0);  // rows that hold a non-fluid cell, ghost rows of a slab included
        bool any_pbc = false;
        for (int x = 0; x < c->NX; x++)
            for (int y = 0; y < c->NY; y++) {
                const uint8_t k = bc->kind_map[(size_t)x * c->NY + y];
                if (k >= bc->n_kinds) return fail(LBM_ERR_ARG, "bc: kind_map[%d,%d] = %d out of range", x, y, k);
                km[(size_t)x * c->pitch + y] = k;
                if (k) {
                    dirty[x] = 1;
                    // Slabs (ghost_x >= 2): the ghost rows mirror the neighbour's rows and carry THEIR kinds — a multi-step
                    // pass recomputes the intermediate states of those rows from the ghost values. They are never stored.
                    const bool slab_ghost = c->gx >= 2 && (x < c->gx || x >= c->NX - c->gx);
                    if (!slab_ghost) cells.push_back(make_int2(x, y));
                    const lbm_kind &kd = bc->kinds[k];
                    if (kd.flags & (LBM_CELL_PBC_IN_SRC | LBM_CELL_PBC_OUT_SRC)) {
                        any_pbc = true;
                        if ((kd.flags & LBM_CELL_PBC_IN_SRC) && x != c->NX - 2) return fail(LBM_ERR_ARG, "bc: PBC_IN_SRC cells must lie on row nx-2");
                        if ((kd.flags & LBM_CELL_PBC_OUT_SRC) && x != 1) return fail(LBM_ERR_ARG, "bc: PBC_OUT_SRC cells must lie on row 1");
                    }
                    if (!slab_ghost && (x < c->gx || x >= c->NX - c->gx || (c->gy && (y == 0 || y == c->NY - 1))))
                        return fail(LBM_ERR_ARG, "bc: ghost cells must be fluid (the reference applies its closures to the interior view)");
                }
            }
        if (any_pbc && (c->gx || c->NX < 4)) return fail(LBM_ERR_ARG, "bc: pressure-periodic boundary needs nx >= 4 and no ghost rows");
        CK(cudaMalloc(&c->kind_map, km.size()));
        CK(cudaMemcpyAsync(c->kind_map, km.data(), km.size(), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMalloc(&c->kinds, bc->n_kinds * sizeof(lbm_kind)));
        CK(cudaMemcpyAsync(c->kinds, bc->kinds, bc->n_kinds * sizeof(lbm_kind), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMalloc(&c->ktab, bc->n_k_rows * 72));
        CK(cudaMemcpyAsync(c->ktab, bc->k_table, bc->n_k_rows * 72, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMalloc(&c->ctab, std::max(bc->n_c_rows, 1) * 72));
        if (bc->n_c_rows) CK(cudaMemcpyAsync(c->ctab, bc->c_table, bc->n_c_rows * 72, cudaMemcpyHostToDevice, c->stream));
        c->n_cells = (int)cells.size();
        if (c->n_cells) {
            CK(cudaMalloc(&c->cells, cells.size() * sizeof(int2)));
            CK(cudaMemcpyAsync(c->cells, cells.data(), cells.size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
            // (same size conditions as fused_ok: small lattices never take the two-step pass)
            if (!any_pbc && c->gx != 1 && !c->gy && c->NY >= 2 * std::max(kFusedThreads, kDeepThreads) && (c->NY % 2) == 0 &&
                (long long)c->NX * c->NY >= kEdgeThreshold)
                if (int rc = plan_strips(c, dirty)) return rc;
        } else if (c->gx >= 2 && std::find(dirty.begin(), dirty.end(), (char)1) != dirty.end() && !any_pbc && !c->gy &&
                   c->NY >= 2 * std::max(kFusedThreads, kDeepThreads) && (c->NY % 2) == 0 && (long long)c->NX * c->NY >= kEdgeThreshold) {
            if (int rc = plan_strips(c, dirty)) return rc;   // only the neighbours' rows hold boundary cells: still a BC slab
        } else {
            c->has_bc = false;
        }
    }
    CK(cudaStreamSynchronize(c->stream));
    return LBM_OK;
}

extern "C" int lbm_create(int device, int nx, int ny, int ghost_x, int ghost_y, const lbm_bc_desc *bc, lbm_ctx **out)
{
    if (!out) return fail(LBM_ERR_ARG, "lbm_create: out is null");
    *out = nullptr;
    if (nx < 1 || ny < 1) return fail(LBM_ERR_ARG, "lbm_create: lattice must be at least 1x1 (got %dx%d)", nx, ny);
    if (ghost_x < 0 || ghost_x > kMaxDepth || (ghost_y & ~1)) return fail(LBM_ERR_ARG, "lbm_create: ghost_x must be 0..%d and ghost_y 0 or 1", kMaxDepth);
    if (ghost_x >= 2 && ghost_y)
        return fail(LBM_ERR_ARG, "lbm_create: %d ghost rows (slabs of the multi-step kernel) exist for lattices without y ghosts only", ghost_x);
    if ((ghost_x && nx < 4 * ghost_x) || (ghost_y && ny < 3)) return fail(LBM_ERR_ARG, "lbm_create: too few interior cells for the ghost ring");
    lbm_ctx *c = new lbm_ctx;
    c->device = device;
    c->NX = nx;
    c->NY = ny;
    c->gx = ghost_x;
    c->gy = ghost_y;
    c->pitch = (ny + 15) & ~15;
    c->plane = (long long)nx * c->pitch;
    if (const char *g = getenv("LBM_GENERIC_KERNEL")) c->force_generic = atoi(g) != 0;
    if (const char *g = getenv("LBM_NO_GRAPHS")) c->use_graphs = atoi(g) == 0;
    if (const char *g = getenv("LBM_NO_FUSED")) c->use_fused = atoi(g) == 0;
    if (const char *g = getenv("LBM_NO_CLUSTER")) c->use_cluster = atoi(g) == 0;
    if (const char *g = getenv("LBM_FUSED_SEG")) c->fused_seg = atoi(g) >= 2 ? atoi(g) : 0;
    if (const char *g = getenv("LBM_FUSED_DEPTH")) c->fused_depth = std::min(kMaxDepth, std::max(2, atoi(g)));
    if (const char *g = getenv("LBM_DEEP2")) c->deep2 = atoi(g) != 0;
    if (const char *t = getenv("LBM_HALO_TIMEOUT_S")) c->timeout_cycles = (long long)(atof(t) * 2e9);
    if (int rc = ctx_build(c, bc)) {
        std::string keep = g_err;
        lbm_destroy(c);
        g_err = keep;
        return rc;
    }
    *out = c;
    return LBM_OK;
}

extern "C" int lbm_set_bc_mode(lbm_ctx *c, int mode)
{
    if (!c || mode < LBM_BC_AUTO || mode > LBM_BC_EDGE) return fail(LBM_ERR_ARG, "lbm_set_bc_mode: bad argument");
    c->bc_mode = mode;
    return LBM_OK;
}

extern "C" int lbm_set_option(lbm_ctx *c, const char *name, int value)
{
    if (!c || !name) return fail(LBM_ERR_ARG, "lbm_set_option: null argument");
    const std::string n(name);
    if (n == "fused")
        c->use_fused = value != 0;
    else if (n == "graphs")
        c->use_graphs = value != 0;
    else if (n == "pdl") {
        c->pdl = value != 0;
        drop_graphs(c);
    }
    else if (n == "generic_kernel")
        c->force_generic = value != 0;
    else if (n == "fused_exact")
        c->fused_exact = value != 0;
    else if (n == "l2_prefetch") {
        if (value < 0 || value > 8) return fail(LBM_ERR_ARG, "lbm_set_option: l2_prefetch must be 0..8 rows");
        c->l2_prefetch = value;
    } else if (n == "streamed") {
        c->use_streamed = value != 0;
    } else if (n == "streamed_chunk_rows") {
        if (value < 0) return fail(LBM_ERR_ARG, "lbm_set_option: streamed_chunk_rows must be >= 0");
        c->streamed_chunk_rows = value;
    } else if (n == "wave_seg") {
        c->wave_seg = value != 0;
    } else if (n == "tail") {
        c->force_tail = value != 0;
    } else if (n == "cluster") {   // 0 = never, 1 = where measured faster than graph replay (default), 2 = wherever the lattice fits
        if (value < 0 || value > 2) return fail(LBM_ERR_ARG, "lbm_set_option: cluster must be 0, 1 or 2");
        c->use_cluster = value != 0;
        c->tune_pick = value == 2 ? 1 : 0;
        c->tune_calls = 0;
        for (auto &t : c->tune) t.pending = false;
    } else if (n == "max_queued_calls") {
        if (value < 0 || value >= lbm_ctx::kCallRing) return fail(LBM_ERR_ARG, "lbm_set_option: max_queued_calls must be 0 (unbounded) .. %d", lbm_ctx::kCallRing - 1);
        c->max_queued_calls = value;
    } else if (n == "fused_depth") {
        if (value < 2 || value > kMaxDepth) return fail(LBM_ERR_ARG, "lbm_set_option: fused_depth must be 2..%d", kMaxDepth);
        c->fused_depth = value;
    } else if (n == "deep2") {
        c->deep2 = value != 0;

    } else if (n == "fused_seg") {
        if (value < 2 && value != 0) return fail(LBM_ERR_ARG, "lbm_set_option: fused_seg must be >= 2, or 0 for the default");
        c->fused_seg = value;
    } else
        return fail(LBM_ERR_ARG, "lbm_set_option: unknown option '%s' (fused, graphs, pdl, generic_kernel, fused_exact, fused_seg, fused_depth, deep2, l2_prefetch, max_queued_calls, cluster, tail, wave_seg, streamed, streamed_chunk_rows)", name);
    return LBM_OK;
}

extern "C" int64_t lbm_device_bytes(const lbm_ctx *c) { return c ? (int64_t)c->arena_bytes + c->stage_cells * 96 : 0; }
extern "C" void *lbm_stream(lbm_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int64_t lbm_time(const lbm_ctx *c) { return c ? c->t : -1; }
extern "C" int64_t lbm_launch_count(const lbm_ctx *c) { return c ? c->launches : 0; }

static bool use_mask(const lbm_ctx *c)
{
    if (!c->has_bc) return false;
    if (c->bc_mode == LBM_BC_MASK) return true;
    if (c->bc_mode == LBM_BC_EDGE) return false;
    return (long long)c->NX * c->NY < kEdgeThreshold;
}

static void fill_common(const lbm_ctx *c, StepParams &P, int src_buf, int dst_buf, double omega)
{
    memset(&P, 0, sizeof P);
    P.src = c->S[src_buf];
    P.dst = c->S[dst_buf];
    P.plane = c->plane;
    P.dplane = c->plane;
    P.pitch = c->pitch;
    P.NX = c->NX;
    P.NY = c->NY;
    P.gx = c->gx;
    P.gy = c->gy;
    P.omega = omega;
    P.omega_last = omega;
    P.kind_map = c->has_bc ? c->kind_map : nullptr;
    P.kinds = c->kinds;
    P.ktab = c->ktab;
    P.ctab = c->ctab;
    // outlet side buffers alternate with the S buffers: the kernel that reads S[b] reads outbuf[b]
    P.out_cur = c->outbuf[src_buf];
    P.out_next = c->outbuf[dst_buf];
    P.rho_in = c->rho_in;
    P.rho_out = c->rho_out;
    P.px = -1;
    P.py = -1;
    P.snap_row = c->snap_row;
    P.no_snap = c->skip_snap ? 1 : 0;
    P.eager_progress = c->eager_progress ? 1 : 0;
    P.snap_col = c->snap_col;
    P.cells = c->cells;
    P.n_cells = c->n_cells;
    P.done_counter = c->done_counter;
    P.timeout_cycles = c->timeout_cycles;
    P.err_flag = c->err_flag;
    P.err_host = c->err_host;
    P.flag_in = c->flags_in;
}

// Probe recording of a launch that reads the state whose time is tcount[tc_in] and produces tcount[tc_out]
static void set_probe(const lbm_ctx *c, StepParams &P, int tc_in, int tc_out)
{
    if (!c->probe || c->px < 0) return;
    P.px = c->px;
    P.py = c->py;
    P.probe = c->probe;
    P.progress = c->progress;
    P.probe_cap = c->probe_cap;
    P.tc_in = c->tcount + tc_in;
    P.tc_out = c->tcount + tc_out;
}

// arena offsets of a peer (its S[1] offset depends on ITS plane size)
static size_t peer_off_S(const Peer &pr, int buf)
{
    const size_t sbytes = (size_t)9 * pr.nx * pr.pitch * 8;
    return buf == 0 ? 0 : align_up(sbytes, 256);
}
static size_t peer_off_flags(const Peer &pr)
{
    const size_t sbytes = (size_t)9 * pr.nx * pr.pitch * 8;
    return 2 * align_up(sbytes, 256);
}

static void fill_halo(const lbm_ctx *c, StepParams &P, int dst_buf, bool flags)
{
    for (int s = 0; s < 9; s++) {
        P.halo[s].base = nullptr;
        P.flag_out[s] = nullptr;
        const Peer &pr = c->peer[s];
        if (!c->halo_ready || !pr.connected) continue;
        P.halo[s].base = (double *)(pr.arena + peer_off_S(pr, dst_buf));
        P.halo[s].nx = pr.nx;
        P.halo[s].ny = pr.ny;
        P.halo[s].pitch = pr.pitch;
        P.halo[s].plane = (long long)pr.nx * pr.pitch;
        if (flags && pr.remote) {
            // I am the neighbour's slot (8 - s): dx,dy mirrored
            P.flag_out[s] = (unsigned *)(pr.arena + peer_off_flags(pr)) + (8 - s);
        }
    }
}

static int block_size(int w) { return w >= 256 ? 256 : std::max(32, (w + 31) & ~31); }

// pdl: launch with programmatic stream serialization (see pdl_release / pdl_wait); captured into a graph this becomes a
// programmatic dependency edge
template <class... Args>
static cudaError_t launch_kernel(void (*kernel)(Args...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, const StepParams &P)
{
    if (!pdl) {
        kernel<<<grid, block, 0, st>>>(P);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, P);
}

template <bool MASK, bool HALO, bool FINAL, bool LIST>
static cudaError_t launch(const StepParams &P, int blocks, int threads, cudaStream_t st, bool pdl = false)
{
    return launch_kernel(k_step<MASK, HALO, FINAL, LIST>, dim3(blocks), dim3(threads), st, pdl, P);
}

// launch-bound lattices without remote neighbours chain their step kernels with programmatic dependent launches
static bool use_pdl(const lbm_ctx *c) { return c->pdl && !c->any_remote && (long long)c->NX * c->NY < kEdgeThreshold; }

static int rows_launch(lbm_ctx *c, StepParams P, int row0a, int na, int row0b, int nb, bool mask, bool halo, cudaStream_t st)
{
    if (na + nb <= 0) return LBM_OK;
    P.row0a = row0a;
    P.na = na;
    P.row0b = row0b;
    P.y0 = c->gy;
    P.y1 = c->NY - c->gy;
    const int bs = block_size(P.y1 - P.y0);
    P.bpr = (P.y1 - P.y0 + bs - 1) / bs;
    const int blocks = (na + nb) * P.bpr;
    P.n_blocks = blocks;
    cudaError_t e;
    if (!mask && !halo && nb == 0 && !c->gy && (c->NY % 2) == 0 && c->NY >= 64 && na <= 65535 && !c->force_generic) {
        // bandwidth kernel: two cells per thread, 2-D grid
        const int pairs = c->NY / 2, pbs = pairs >= 128 ? 128 : std::max(32, (pairs + 31) & ~31);
        dim3 grid((pairs + pbs - 1) / pbs, na);
        // (the pair kernel has no prologue to overlap: early-launched blocks only take slots away below ~2^17 cells,
        //  profiles/r01e_small_lattices_pdl.txt)
        const long long cells = (long long)c->NX * c->NY;
        e = launch_kernel(cells <= (1 << 19) ? k_step_pair<true> : k_step_pair<false>, grid, dim3(pbs), st, use_pdl(c) && cells >= (1 << 17), P);
    } else if (mask)
        e = halo ? launch<true, true, false, false>(P, blocks, bs, st, use_pdl(c)) : launch<true, false, false, false>(P, blocks, bs, st, use_pdl(c));
    else
        e = halo ? launch<false, true, false, false>(P, blocks, bs, st, use_pdl(c)) : launch<false, false, false, false>(P, blocks, bs, st, use_pdl(c));
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "step kernel launch failed: %s", cudaGetErrorString(e));
    c->launches++;
    return LBM_OK;
}

// One reference time step: S[src] -> S[src^1]. `t_new` is the time index of the state being produced.
static int one_step(lbm_ctx *c, int src, double omega, long long t_new)
{
    const int dst = src ^ 1;
    StepParams P;
    fill_common(c, P, src, dst, omega);
    const bool mask = use_mask(c);
    // mask-free kernel + thin fix-up kernel over the non-fluid cells (none on a slab whose only boundary cells sit on its ghost rows)
    const bool fix = c->has_bc && !mask && c->n_cells > 0;
    const bool halo = c->halo_ready;
    const bool remote = halo && c->any_remote;
    fill_halo(c, P, dst, true);
    (void)t_new;
    set_probe(c, P, src, dst);
    const int xlo = c->gx, xhi = c->NX - c->gx;   // interior rows [xlo, xhi)
    // Flag protocol (remote neighbours only): a kernel that reads my ghosts / writes a neighbour's ghosts first
    // waits until every neighbour has published epoch E (= it finished producing its S_t, hence finished reading
    // the buffer I am about to overwrite), and the LAST kernel of the step that stores ghosts publishes E+1.
    const unsigned E = c->halo_epoch;
    const int g = c->gx;
    const bool split = halo && g && !c->gy && !fix && (xhi - xlo) >= 4 * g && (long long)c->NX * c->NY >= kEdgeThreshold;
    if (!split) {
        if (remote) {
            P.wait_value = E;
            P.signal_value = fix ? 0 : E + 1;
        }
        if (int rc = rows_launch(c, P, xlo, xhi - xlo, 0, 0, mask, halo, c->stream)) return rc;
    } else {
        // 1-D slabs: the two edge rows (the only readers of ghost rows and the only writers of neighbour ghosts)
        // run on the high-priority stream; the interior overlaps with the NVLink stores they generate.
        CK(cudaEventRecord(c->ev_main, c->stream));
        CK(cudaStreamWaitEvent(c->stream_edge, c->ev_main, 0));   // everything queued so far (previous step's join)
        StepParams Pe = P;
        if (remote) {
            Pe.wait_value = E;
            Pe.signal_value = E + 1;
        }
        if (int rc = rows_launch(c, Pe, xlo, g, xhi - g, g, mask, true, c->stream_edge)) return rc;
        CK(cudaEventRecord(c->ev_edge, c->stream_edge));
        StepParams Pi = P;
        for (int s = 0; s < 9; s++) Pi.halo[s].base = nullptr;
        if (int rc = rows_launch(c, Pi, xlo + g, xhi - xlo - 2 * g, 0, 0, mask, false, c->stream)) return rc;
        CK(cudaStreamWaitEvent(c->stream, c->ev_edge, 0));        // join: the next step needs both
    }
    if (fix) {
        StepParams L = P;
        L.wait_value = 0;
        L.signal_value = remote ? E + 1 : 0;
        const int blocks = (c->n_cells + 127) / 128;
        L.n_blocks = blocks;
        cudaError_t e = halo ? launch<true, true, false, true>(L, blocks, 128, c->stream, use_pdl(c)) : launch<true, false, false, true>(L, blocks, 128, c->stream, use_pdl(c));
        if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "edge kernel launch failed: %s", cudaGetErrorString(e));
        c->launches++;
    }
    if (remote) c->halo_epoch++;
    return LBM_OK;
}

// Two reference time steps in one pass: S[src] (time t) -> S[src^1] (time t+2). Fluid lattices, no y ghosts.

static bool fused_ok(const lbm_ctx *c)
{
    return c->use_fused && (!c->has_bc || c->fused_bc) && !c->gy && c->gx != 1 && c->NY >= 2 * std::max(kFusedThreads, kDeepThreads) && (c->NY % 2) == 0 &&
           (c->NX - 2 * c->gx) >= 8 && (long long)c->NX * c->NY >= kEdgeThreshold && (!c->gx || c->halo_ready);
}

// May a call END on a multi-step pass? Yes: results of time t are rebuilt from S_{t-d} by re-running the pass with its
// last level in FINAL mode (fluid rows), or from the strip windows, which keep S_{t-1} of the boundary rows (lattices
// with boundary cells); option "tail" = 1 still ends every call with a one-step launch (A/B, tests).
static bool tail_free(const lbm_ctx *) { return true; }

// Deepest pass this lattice can take: lattices with boundary cells stay on two steps (strip windows), slabs cannot look
// further than their ghost rows.
static int max_depth(const lbm_ctx *c)
{
    const int d = c->gx ? std::min(c->gx, c->fused_depth) : c->fused_depth;
    return c->has_bc ? std::min(d, c->bc_depth) : d;   // strips are planned for bc_depth steps (or fewer) per pass
}

// Output rows per thread block of the multi-step kernels. Every block recomputes D-1 intermediate rows per level at each
// end of its segment (redundant work ~ 2(D-1)/seg), but a lattice of a few million cells needs short segments to fill the
// SMs several times over. Measured in round 1 (tools/fused_sweep.py, profiles/r01c_fused_sweep.txt): the best segment is
// the longest one that still gives ~2000 blocks — 8 rows at 1536^2, 16 at 3072^2, 32 at 4096^2, 64+ from 8192^2 on.
// Among the long segments the one whose block count fills whole WAVES of resident blocks wins: 16384^2 with 256-row
// segments is 67 x 64 = 4288 blocks = 14.5 waves of 296 (the last wave half empty: 3.4 % of the launch idle), with
// 310-row segments 67 x 53 = 3551 blocks = 12.0 waves.
static int pick_seg(const lbm_ctx *c, int rows, int depth = 2)
{
    if (c->fused_seg) return c->fused_seg;
    const bool deep = depth > 2 || c->deep2;
    const int T = deep ? kDeepThreads : kFusedThreads;
    const long long strips = (c->NY + (2 * T - 4 * (depth - 1)) - 1) / (2 * T - 4 * (depth - 1));
    int seg = 256;
    while (seg > 8 && strips * ((rows + seg - 1) / seg) < 2000) seg /= 2;
    if (seg < 128 || !c->wave_seg) return seg;
    const long long slots = (long long)c->n_sm * (deep ? (depth == 2 ? 3 : 2) * (kDeepThreads <= 64 ? 128 / kDeepThreads : 1) : 4);
    double best_eff = 0;
    int best = seg;
    for (int ns = (rows + 511) / 512; ns <= (rows + 127) / 128; ns++) {      // segments of 128 .. 512 rows
        const int sg = (rows + ns - 1) / ns;
        const long long blocks = strips * ((rows + sg - 1) / sg);
        const double waves = (double)blocks / slots, eff = waves / std::ceil(waves);
        if (eff > best_eff + 1e-3) {        // (ties: the fewest, longest segments)
            best_eff = eff;
            best = sg;
        }
    }
    return best;
}

template <bool HALO>
static int fused_launch(lbm_ctx *c, StepParams P, int row0a, int na, int row0b, int nb, int seg, cudaStream_t st, int depth = 2,
                        bool force_deep = false)
{
    if (na + nb <= 0) return LBM_OK;
    constexpr int T = kFusedThreads;
    if (depth > 2 || c->deep2 || force_deep) {   // k_stepNx
        P.row0a = row0a;
        P.na = na;
        P.row0b = row0b;
        P.nb = nb;
        P.seg = seg;
        P.pf = c->l2_prefetch;
        const int W = deep_width(depth);
        dim3 grid((c->NY + W - 1) / W, (na + seg - 1) / seg + (nb + seg - 1) / seg);
        P.n_blocks = (int)(grid.x * grid.y);
        deep_kernel(depth, HALO, P.probe != nullptr, false)<<<grid, kDeepThreads, deep_smem(depth), st>>>(P);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "%d-step kernel launch failed: %s", depth, cudaGetErrorString(e));
        c->launches++;
        return LBM_OK;
    }
    const size_t smem = (size_t)4 * 6 * 2 * T * sizeof(double);
    // (the opt-in to > 48 KB of dynamic shared memory is done once per device in ctx_build: cudaFuncSetAttribute
    //  may wait for the device, and a neighbour's kernel may be spinning on a flag this rank has yet to publish)
    P.row0a = row0a;
    P.na = na;
    P.row0b = row0b;
    P.nb = nb;
    P.seg = seg;
    P.pf = c->l2_prefetch;
    dim3 grid((c->NY + 2 * T - 5) / (2 * T - 4), (na + seg - 1) / seg + (nb + seg - 1) / seg);
    P.n_blocks = (int)(grid.x * grid.y);
    if (P.probe)
        k_step2x<T, HALO, true><<<grid, T, smem, st>>>(P);
    else
        k_step2x<T, HALO, false><<<grid, T, smem, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "two-step kernel launch failed: %s", cudaGetErrorString(e));
    c->launches++;
    return LBM_OK;
}

// Lattice with boundary cells (plan_strips): the multi-step kernel on the clean rows, `depth` one-step mask launches per
// strip through its windows: S_t rows [a-d, b+d) -> ... -> the last window (rows [a-1, b+1) of S_{t+d-1}) -> S_{t+d} rows [a, b).
// All launches read S[src] (and the strip windows) and write disjoint rows of S[dst]: plain stream order.
// Slabs (gx >= depth): the strips next to the slab edges are the only launches that touch ghost rows. Their first launch
// (ghost rows included) waits for the neighbours' previous pass; their last launch stores the rows within gx of the
// edge into the neighbours' ghost rows as well, and the last of them publishes the pass.
static int two_steps_bc(lbm_ctx *c, const StepParams &P0, int src, int dst, int depth)
{
    const int NX = c->NX;
    const bool probe = c->probe && c->px >= 0;
    const bool slab = c->gx > 0, remote = slab && c->any_remote;
    const unsigned E = c->halo_epoch;
    for (size_t i = 0; i < c->clean.size(); i += 2) {
        StepParams P = P0;
        const auto ra = c->clean[i], rb = i + 1 < c->clean.size() ? c->clean[i + 1] : std::make_pair(0, 0);
        if (probe && ((c->px >= ra.first && c->px < ra.second) || (c->px >= rb.first && c->px < rb.second))) set_probe(c, P, src, dst);
        // (a redo with a new omega for the last collision needs the kernel that knows omega_last: k_stepNx)
        if (int rc = fused_launch<false>(c, P, ra.first, ra.second - ra.first, rb.first, rb.second - rb.first, pick_seg(c, c->NX, depth), c->stream, depth,
                                         P0.omega_last != P0.omega)) return rc;
    }
    int last_edge = -1;
    for (size_t i = 0; i < c->strips.size(); i++)
        if (slab && (c->strips[i].a == c->gx || c->strips[i].b == NX - c->gx)) last_edge = (int)i;
    for (size_t i = 0; i < c->strips.size(); i++) {
        const auto &s = c->strips[i];
        const bool edge = slab && (s.a == c->gx || s.b == NX - c->gx);
        const bool probe_here = probe && ((c->px >= s.a && c->px < s.b) || (c->px + NX >= s.a && c->px + NX < s.b));
        // rows [lo, hi) of the lattice, unwrapped, as (at most) two wrapped ranges
        auto launch_rows = [&](StepParams &P, int lo, int hi) {
            if (lo < 0) return rows_launch(c, P, lo + NX, -lo, 0, hi, true, edge, c->stream);
            if (hi > NX) return rows_launch(c, P, lo, NX - lo, 0, hi - NX, true, edge, c->stream);
            return rows_launch(c, P, lo, hi - lo, 0, 0, true, edge, c->stream);
        };
        // stage k = 1 .. depth reads stage k-1's buffer (0: S[src]) and writes stage k's (depth: S[dst]); the windows of a
        // three-step pass are buf1 (rows a-2 ..) then buf (rows a-1 ..), a two-step pass uses buf only
        for (int k = 1; k <= depth; k++) {
            const int m_in = depth - (k - 1), m_out = depth - k;           // margins: rows [a - m, b + m)
            StepParams P = P0;
            auto window = [&](int m, double *&ptr, long long &plane, int &base, int &ob, int &tc) {
                ptr = m == 1 ? s.buf : s.buf1;
                plane = (long long)(s.b - s.a + 2 * m) * c->pitch;
                base = (s.a + NX - m) % NX;
                ob = m == 1 ? 2 : 3;
                tc = m == 1 ? 2 : 3;
            };
            int tc_in = src, tc_out = dst;
            if (k > 1) {
                double *ptr; long long plane; int base, ob;
                window(m_in, ptr, plane, base, ob, tc_in);
                P.src = ptr;
                P.plane = plane;
                P.sbase = base;
                P.out_cur = c->outbuf[ob];
            }
            if (k < depth) {
                double *ptr; long long plane; int base, ob;
                window(m_out, ptr, plane, base, ob, tc_out);
                P.dst = ptr;
                P.dplane = plane;
                P.dbase = base;
                P.out_next = c->outbuf[ob];
            } else {
                P.omega = P0.omega_last;     // the pass's LAST collision (differs when the caller changed omega)
            }
            P.no_snap = 1;
            if (edge) {   // ghost rows: wait before the first stage, store into the neighbours' and publish in the last
                fill_halo(c, P, dst, true);
                if (k < depth)
                    for (int q = 0; q < 9; q++) P.halo[q].base = nullptr;
                if (remote && k == 1) P.wait_value = E;
                if (remote && k == depth && (int)i == last_edge) P.signal_value = E + 1;
            }
            if (probe_here) set_probe(c, P, tc_in, tc_out);
            if (int rc = launch_rows(P, s.a - m_out, s.b + m_out)) return rc;
        }
    }
    if (remote) c->halo_epoch++;
    return LBM_OK;
}

static int two_steps(lbm_ctx *c, int src, double omega, int depth = 2, double omega_last = 0.0)
{
    const int dst = src ^ 1;
    StepParams P;
    fill_common(c, P, src, dst, omega);
    if (omega_last != 0.0) P.omega_last = omega_last;   // redo of a pass whose last collision must use the caller's new omega
    if (c->has_bc) return two_steps_bc(c, P, src, dst, depth);
    set_probe(c, P, src, dst);
    const int g = c->gx, xlo = g, xhi = c->NX - g;
    if (!g) return fused_launch<false>(c, P, xlo, xhi - xlo, 0, 0, pick_seg(c, xhi - xlo, depth), c->stream, depth, P.omega_last != P.omega);
    // two-row slabs: the 2 + 2 edge rows (readers of the ghost rows, writers of the neighbours') first on the
    // high-priority stream with the flag handshake, the interior overlaps with their NVLink stores
    const bool remote = c->any_remote;
    fill_halo(c, P, dst, true);
    const unsigned E = c->halo_epoch;
    CK(cudaEventRecord(c->ev_main, c->stream));
    CK(cudaStreamWaitEvent(c->stream_edge, c->ev_main, 0));
    StepParams Pe = P;
    if (remote) {
        Pe.wait_value = E;
        Pe.signal_value = E + 1;
    }
    if (int rc = fused_launch<true>(c, Pe, xlo, g, xhi - g, g, g, c->stream_edge, depth, P.omega_last != P.omega)) return rc;
    CK(cudaEventRecord(c->ev_edge, c->stream_edge));
    StepParams Pi = P;
    for (int s = 0; s < 9; s++) Pi.halo[s].base = nullptr;
    if (int rc = fused_launch<false>(c, Pi, xlo + g, xhi - xlo - 2 * g, 0, 0, pick_seg(c, xhi - xlo - 2 * g, depth), c->stream, depth,
                                     P.omega_last != P.omega)) return rc;
    CK(cudaStreamWaitEvent(c->stream, c->ev_edge, 0));
    if (remote) c->halo_epoch++;
    return LBM_OK;
}

static int first_collide(lbm_ctx *c, InitParams &Q, int x0, int nrows)
{
    Q.x0 = x0;
    Q.nrows = nrows;
    const int bs = block_size(c->NY);
    Q.S.bpr = (c->NY + bs - 1) / bs;
    const int blocks = nrows * Q.S.bpr;
    if (c->halo_ready)
        k_first_collide<true><<<blocks, bs, 0, c->stream>>>(Q);
    else
        k_first_collide<false><<<blocks, bs, 0, c->stream>>>(Q);
    CK(cudaGetLastError());
    c->launches++;
    return LBM_OK;
}

static int begin_load(lbm_ctx *c, double omega, InitParams &Q)
{
    if (!(omega > 0.0 && omega < 2.0)) return fail(LBM_ERR_ARG, "omega must satisfy 0 < omega < 2 (got %.17g)", omega);
    if ((c->gx || c->gy) && !c->halo_ready)
        return fail(LBM_ERR_STATE, "a lattice with a ghost ring needs its halo neighbours connected (lbm_halo_connect/finalize) before loading");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->stream_edge));
    CK(cudaMemsetAsync(c->err_flag, 0, 4, c->stream));
    CK(cudaMemsetAsync(c->done_counter, 0, 8, c->stream));
    *c->err_host = 0;
    memset(&Q, 0, sizeof Q);
    c->cur = 0;
    fill_common(c, Q.S, 1, 0, omega);   // "dst" = S[0]; outlet buffer written = outbuf[0], read by the first step
    fill_halo(c, Q.S, 0, false);
    return LBM_OK;
}

static int end_load(lbm_ctx *c, double omega)
{
    CK(cudaMemsetAsync(c->tcount, 0, 32, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *c->progress = 0;
    c->loaded = true;
    c->t = 0;
    c->last_depth = 0;
    c->omega = omega;
    // Ghost stores of the first collision are ordered against the first step by a host-side barrier the caller
    // performs (python: process-group barrier after upload); flags restart from a common epoch.
    return LBM_OK;
}

extern "C" int lbm_upload(lbm_ctx *c, const double *f, const double *rho, const double *u, double omega)
{
    if (!c || !f || !rho || !u) return fail(LBM_ERR_ARG, "lbm_upload: null pointer");
    InitParams Q;
    if (int rc = begin_load(c, omega, Q)) return rc;
    const long long chunk_rows = c->stage_cells / c->NY;
    for (int x0 = 0; x0 < c->NX; x0 += (int)chunk_rows) {
        const int nr = (int)std::min<long long>(chunk_rows, c->NX - x0);
        const size_t n = (size_t)nr * c->NY, o = (size_t)x0 * c->NY;
        CK(cudaMemcpyAsync(c->stage_f, f + o * 9, n * 72, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->stage_rho, rho + o, n * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->stage_u, u + o * 2, n * 16, cudaMemcpyHostToDevice, c->stream));
        Q.in_f = c->stage_f;
        Q.in_rho = c->stage_rho;
        Q.in_u = c->stage_u;
        if (int rc = first_collide(c, Q, x0, nr)) return rc;
        CK(cudaStreamSynchronize(c->stream));
    }
    return end_load(c, omega);
}

extern "C" int lbm_init_equilibrium(lbm_ctx *c, const double *rho_x, const double *ux_y, double rho0, double ux0, double uy0,
                                    double omega)
{
    if (!c) return fail(LBM_ERR_ARG, "lbm_init_equilibrium: null context");
    InitParams Q;
    if (int rc = begin_load(c, omega, Q)) return rc;
    DevBuf dr, du;
    if (rho_x) {
        CK(dr.alloc((size_t)c->NX * 8));
        CK(cudaMemcpyAsync(dr.p, rho_x, (size_t)c->NX * 8, cudaMemcpyHostToDevice, c->stream));
        Q.rho_x = dr.as<double>();
    }
    if (ux_y) {
        CK(du.alloc((size_t)c->NY * 8));
        CK(cudaMemcpyAsync(du.p, ux_y, (size_t)c->NY * 8, cudaMemcpyHostToDevice, c->stream));
        Q.ux_y = du.as<double>();
    }
    Q.rho0 = rho0;
    Q.ux0 = ux0;
    Q.uy0 = uy0;
    if (int rc = first_collide(c, Q, 0, c->NX)) return rc;
    return end_load(c, omega);
}

// ---- thread-block cluster kernel for lattices that fit in distributed shared memory -------------------------------
typedef void (*cluster_fn)(const ClusterParams);
template <bool MASK, int M>
static cluster_fn cluster_kernel_t(int threads)   // the register budget follows the block size (512: 128, 768: 80, 1024: 64)
{
    return threads <= 512 ? k_cluster_steps<MASK, M, 512> : (threads <= 768 ? k_cluster_steps<MASK, M, 768> : k_cluster_steps<MASK, M, 1024>);
}
static cluster_fn cluster_kernel(bool mask, int m, int threads)
{
    if (mask) return m == 1 ? cluster_kernel_t<true, 1>(threads) : cluster_kernel_t<true, 2>(threads);
    return m == 1 ? cluster_kernel_t<false, 1>(threads) : cluster_kernel_t<false, 2>(threads);
}

// Decides once per context whether (and how) the cluster kernel can run this lattice: no ghost ring, both A/B buffers of
// ceil(NX / C) rows within one CTA's shared memory, at most two cells per thread; C = 16 (non-portable size) when the
// device can co-schedule such a cluster, else 8.
static void cluster_plan(lbm_ctx *c)
{
    c->cluster_size = -1;
    if (c->gx || c->gy || c->halo_ready) return;
    const long long cells = (long long)c->NX * c->NY;
    if (cells > 16LL * 2048 || c->NX < 2) return;
    const bool mask = c->has_bc;
    const int force_c = getenv("LBM_CLUSTER_SIZE") ? atoi(getenv("LBM_CLUSTER_SIZE")) : 0;
    for (int C : {16, 8, 4, 2, 1}) {
        if (C > c->NX || (force_c && C != force_c) || (!force_c && C < 8)) continue;
        const int R = (c->NX + C - 1) / C;                  // most rows a CTA owns (balanced split, k_cluster_steps)
        const long long per = (long long)R * c->NY;
        const size_t smem = (size_t)2 * 9 * per * sizeof(double);
        if (per > 2048 || smem > 220 * 1024) continue;
        // one cell per thread up to 768 threads (more warps hide the latency of the cell update better than the larger
        // register budget of a smaller block does: Couette 100 x 100 on 16 CTAs 1.77 vs 2.20 us per step), two cells per thread above
        const int mlim = getenv("LBM_CLUSTER_MLIM") ? atoi(getenv("LBM_CLUSTER_MLIM")) : 768;   // (A/B knob)
        const int m = per > mlim ? 2 : 1;
        const int threads = (int)std::min<long long>(1024, ((per + m - 1) / m + 31) / 32 * 32);
        cluster_fn fn = cluster_kernel(mask, m, threads);
        if (cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        if (C > 8 && cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, (const void *)fn, &cfg) != cudaSuccess || n_clusters < 1) {
            cudaGetLastError();
            continue;
        }
        c->cluster_size = C;
        c->cluster_rows = R;
        c->cluster_threads = threads;
        c->cluster_m = m;
        c->cluster_smem = smem;
        return;
    }
}

static bool cluster_ok(lbm_ctx *c)
{
    if (!c->use_cluster) return false;
    if (c->cluster_size == 0) cluster_plan(c);
    return c->cluster_size > 0;
}

// n time steps in one launch: S[cur] (time t) -> S[cur ^ (n & 1)] (time t+n), the other buffer receives time t+n-1
static int cluster_steps(lbm_ctx *c, double omega, int n)
{
    ClusterParams Q;
    memset(&Q, 0, sizeof Q);
    fill_common(c, Q.S, c->cur, c->cur ^ 1, omega);
    set_probe(c, Q.S, c->cur, c->cur ^ 1);
    Q.n_steps = n;
    Q.R = c->cluster_rows;
    Q.n_cells = c->cluster_rows * c->NY;
    Q.ob[0] = c->outbuf[c->cur];
    Q.ob[1] = c->outbuf[c->cur ^ 1];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c->cluster_size);
    cfg.blockDim = dim3(c->cluster_threads);
    cfg.dynamicSmemBytes = c->cluster_smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c->cluster_size;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_kernel(c->has_bc, c->cluster_m, c->cluster_threads), Q);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "cluster kernel launch failed: %s", cudaGetErrorString(e));
    c->launches++;
    return LBM_OK;
}

// ---- CUDA graphs for launch-bound lattices ----------------------------------------------------------------
// Configs 1-4 of BASELINE.json are <= 77 k cells: a step is 2-5 us of GPU work behind ~2 us of launch gap. A graph of
// kGraphSteps captured steps is replayed instead; kGraphSteps is even, so the A/B parity is the same before and after.
// (Tried and dropped: all steps in ONE cooperative launch with a grid barrier — atomic counter + generation word —
//  after every step. Bit-exact, but the barrier costs more than a kernel boundary inside a graph: +1.1 us per step on
//  every lattice from 100x50 to 1000x1000, profiles/r01e_persistent_vs_graph.txt.)
static const int kGraphSteps = 32;

static void drop_graphs(lbm_ctx *c)
{
    for (auto &g : c->graphs) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}

static int graph_for(lbm_ctx *c, double omega, bool snap_last, lbm_ctx::GraphEntry **out)
{
    const int mode = c->bc_mode;
    for (auto &g : c->graphs)
        if (g.omega == omega && g.parity == c->cur && g.mode == mode && g.snap_last == snap_last && g.probe == (const void *)c->probe) {
            *out = &g;
            return LBM_OK;
        }
    if (c->graphs.size() >= 8) drop_graphs(c);
    cudaGraph_t graph = nullptr;
    const long long l0 = c->launches;
    CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = LBM_OK, src = c->cur;
    for (int i = 0; i < kGraphSteps && rc == LBM_OK; i++, src ^= 1) {
        c->skip_snap = !(snap_last && i == kGraphSteps - 1);
        rc = one_step(c, src, omega, 0);
    }
    c->skip_snap = false;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    const long long per_graph = c->launches - l0;
    c->launches = l0;
    if (rc != LBM_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(e));
    c->graphs.push_back({omega, c->cur, mode, snap_last, (const void *)c->probe, exec, per_graph});
    *out = &c->graphs.back();
    return LBM_OK;
}

// Reads back the timing samples of the first calls (without blocking) and decides once all four are in.
static void tune_resolve(lbm_ctx *c)
{
    if (c->tune_pick || c->tune_calls < 4) return;
    double best[3] = {0, 1e30, 1e30};   // us per step, the better of the two samples of a path (the first one may include
    for (auto &t : c->tune) {           // graph capture / module load on the host while the device idles)
        if (!t.pending || cudaEventQuery(t.b) != cudaSuccess) {
            cudaGetLastError();
            return;   // not finished yet: ask again at the next call
        }
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
        if (t.steps > 0) best[t.path] = std::min(best[t.path], 1e3 * ms / t.steps);
    }
    c->tune_pick = best[1] <= best[2] ? 1 : 2;
}

// A halo wait that timed out is reported by every later call that looks at the lattice (sticky until the next load):
// the kernels after the failed wait have not computed anything (halo_wait), so the state is unusable.
static int async_error(lbm_ctx *c, const char *who)
{
    const unsigned err = c->err_host ? *(volatile unsigned *)c->err_host : 0u;
    if (!err) return LBM_OK;
    return fail(LBM_ERR_TIMEOUT,
                "%s: halo flag wait timed out — a neighbouring rank did not take the same step (waited for step %u of neighbour "
                "slot %u; my step count %u). The lattice state is invalid: load it again on every rank.",
                who, (err >> 8) & 0x7fffff, err & 0xff, c->halo_epoch);
}

extern "C" int lbm_step(lbm_ctx *c, double omega, int n_steps)
{
    if (!c) return fail(LBM_ERR_ARG, "lbm_step: null context");
    if (!(omega > 0.0 && omega < 2.0)) return fail(LBM_ERR_ARG, "omega must satisfy 0 < omega < 2 (got %.17g)", omega);
    if (n_steps < 0) return fail(LBM_ERR_ARG, "lbm_step: n_steps < 0");
    if (!c->loaded) return fail(LBM_ERR_STATE, "lbm_step before lbm_upload / lbm_init_equilibrium");
    if (n_steps == 0) return LBM_OK;
    if (int rc = async_error(c, "lbm_step")) return rc;
    CK(cudaSetDevice(c->device));
    if (c->max_queued_calls > 0 && c->n_calls >= c->max_queued_calls)
        CK(cudaEventSynchronize(c->ev_call[(c->n_calls - c->max_queued_calls) % lbm_ctx::kCallRing]));
    if (omega != c->omega) {
        // S[cur] was collided with the previous call's omega. Redo that collision from the retained S_{t-d}: the last
        // launch again (a one-step launch, or a multi-step pass whose LAST level collides with the new omega).
        if (c->t == 0 || c->last_depth < 1)
            return fail(LBM_ERR_STATE, "omega differs from the one the resident state was collided with and the previous state is not retained: upload again");
        if (c->any_remote) return fail(LBM_ERR_STATE, "changing omega between steps is not supported with remote halo neighbours: upload again");
        if (c->last_depth == 1) {
            if (int rc = one_step(c, c->cur ^ 1, omega, c->t)) return rc;
        } else {
            if (int rc = two_steps(c, c->cur ^ 1, c->omega, c->last_depth, omega)) return rc;
        }
        c->omega = omega;
    }
    int left = n_steps;
    // lattices that fit a cluster's shared memory: all steps of the call in one launch (configs 1-3 of BASELINE.json),
    // unless graph replay was measured to be faster on this lattice
    const bool graph_ok = c->use_graphs && !c->any_remote && (long long)c->NX * c->NY < kEdgeThreshold && left >= 2 * kGraphSteps;
    int tune_slot = -1;
    bool via_cluster = left >= 2 && cluster_ok(c);
    if (via_cluster && graph_ok && c->tune_pick == 0) {
        tune_resolve(c);
        if (c->tune_pick == 0 && c->tune_calls < 4) {
            tune_slot = c->tune_calls++;
            via_cluster = (tune_slot & 1) == 0;          // calls 0, 2 on the cluster kernel, calls 1, 3 on graph replay
            lbm_ctx::TuneSample &ts = c->tune[tune_slot];
            if (!ts.a) {
                CK(cudaEventCreate(&ts.a));
                CK(cudaEventCreate(&ts.b));
            }
            ts.steps = left - left % kGraphSteps * (via_cluster ? 0 : 1);
            ts.path = via_cluster ? 1 : 2;
            CK(cudaEventRecord(ts.a, c->stream));
        }
    } else if (via_cluster && c->tune_pick == 2 && graph_ok) {
        via_cluster = false;
    }
    if (via_cluster) {
        while (left > 0) {
            const int n = std::min(left, 1 << 20);
            if (int rc = cluster_steps(c, omega, n)) return rc;
            c->cur ^= n & 1;
            c->t += n;
            left -= n;
            c->last_depth = 1;
        }
    }
    if (c->use_graphs && !c->any_remote && (long long)c->NX * c->NY < kEdgeThreshold && left >= 2 * kGraphSteps) {
        while (left >= kGraphSteps) {
            lbm_ctx::GraphEntry *g = nullptr;
            if (int rc = graph_for(c, omega, left == kGraphSteps, &g)) return rc;   // (the entry may move: look it up per replay)
            CK(cudaGraphLaunch(g->exec, c->stream));
            c->launches += g->launches;
            c->t += kGraphSteps;
            left -= kGraphSteps;
            c->last_depth = 1;
        }
    }
    if (tune_slot >= 0) {
        CK(cudaEventRecord(c->tune[tune_slot].b, c->stream));
        c->tune[tune_slot].pending = true;
    }
    // Bandwidth-bound fluid lattices advance two steps per pass; the call always ENDS with a one-step launch so that
    // the other buffer holds S_{t-1}, from which results are materialised and a changed omega is redone.
    // Bandwidth-bound lattices advance D steps per pass. Fluid lattices may end a call on a pass (results are then
    // materialised by re-running it in FINAL mode, a changed omega redoes its last level); lattices with boundary
    // cells end every call with a one-step launch so that the other buffer holds S_{t-1}.
    if (fused_ok(c)) {
        const int D = max_depth(c);
        const bool no_tail = (tail_free(c) && !c->force_tail) || c->fused_exact;
        while (true) {
            const int d = std::min(D, no_tail ? left : left - 1);
            if (d < 2) break;
            if (int rc = two_steps(c, c->cur, omega, d)) return rc;
            c->cur ^= 1;
            c->t += d;
            left -= d;
            c->last_depth = d;
        }
    }
    for (int i = 0; i < left; i++) {
        c->skip_snap = i + 1 < left;
        c->eager_progress = n_steps == 1;
        const int rc = one_step(c, c->cur, omega, c->t + 1);
        c->skip_snap = false;
        c->eager_progress = false;
        if (rc) return rc;
        c->cur ^= 1;
        c->t++;
        c->last_depth = 1;
    }
    CK(cudaEventRecord(c->ev_call[c->n_calls % lbm_ctx::kCallRing], c->stream));
    c->n_calls++;
    return LBM_OK;
}

extern "C" int lbm_sync(lbm_ctx *c)
{
    if (!c) return fail(LBM_ERR_ARG, "lbm_sync: null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream_edge));
    CK(cudaStreamSynchronize(c->stream));
    return async_error(c, "lbm_sync");
}

// f_post / rho / u of time t are stream+BC+moments of S_{t-1}, which the A/B scheme still holds in S[cur^1].
// to_host: copy to the host pointers; else either reduce (min/max) or, with dev_rho / dev_u set, write the packed fields
// straight into device memory (history slots) without any copy or synchronisation
static int materialize_rows(lbm_ctx *c, int x0, int x1, int y0, int y1, double *f, double *rho, double *u, bool to_host,
                            double *dev_rho = nullptr, double *dev_u = nullptr)
{
    StepParams P;
    fill_common(c, P, c->cur ^ 1, c->cur, c->omega);
    P.use_snap = c->halo_ready && c->snap_row ? 1 : 0;
    const int w = y1 - y0;
    const long long chunk_rows = std::max<long long>(1, c->stage_cells / w);
    for (int xa = x0; xa < x1; xa += (int)chunk_rows) {
        const int nr = (int)std::min<long long>(chunk_rows, x1 - xa);
        P.row0a = xa;
        P.na = nr;
        P.y0 = y0;
        P.y1 = y1;
        const int bs = block_size(w);
        P.bpr = (w + bs - 1) / bs;
        P.ox0 = xa;
        P.oy0 = y0;
        P.ow = w;
        P.o_f = f ? c->stage_f : nullptr;
        P.o_rho = rho || !to_host ? c->stage_rho : nullptr;
        P.o_u = u || !to_host ? c->stage_u : nullptr;
        if (dev_rho || dev_u) {
            P.o_f = nullptr;
            P.o_rho = dev_rho ? dev_rho + (size_t)(xa - x0) * w : nullptr;
            P.o_u = dev_u ? dev_u + (size_t)(xa - x0) * w * 2 : nullptr;
        }
        const int blocks = nr * P.bpr;
        cudaError_t e = cudaSuccess;
        // the call ended on a multi-step pass: S[cur^1] is S_{t-d}. Re-run the pass over rows [lo, lo + n) with its last
        // level in FINAL mode (f_post / rho / u of time t instead of the collision), column strips that meet [y0, y1) only
        auto deep_final = [&](StepParams Q, int lo, int n, int d) {
            const int W = deep_width(d);
            Q.use_snap = 0;
            Q.row0a = lo;
            Q.na = n;
            Q.row0b = 0;
            Q.nb = 0;
            Q.seg = std::min(n, 64);
            Q.pf = 0;
            Q.strip0 = y0 / W;
            dim3 grid((y1 + W - 1) / W - Q.strip0, (n + Q.seg - 1) / Q.seg);
            deep_kernel(d, false, false, true)<<<grid, kDeepThreads, deep_smem(d), c->stream>>>(Q);
            c->launches++;
            return cudaGetLastError();
        };
        if (c->last_depth > 1 && c->has_bc) {
            // lattice with boundary cells after a multi-step pass: clean rows as above; strip rows from their last WINDOW, which
            // still holds S_{t-1} of the strip (+ one row each side) — one FINAL mask launch
            auto owner = [&](int x) -> const lbm_ctx::Strip * {   // the strip that holds lattice row x (strips may wrap), or null
                for (const auto &st : c->strips)
                    if ((x >= st.a && x < st.b) || (x + c->NX >= st.a && x + c->NX < st.b)) return &st;
                return nullptr;
            };
            for (int lo = xa; lo < xa + nr && e == cudaSuccess;) {
                const lbm_ctx::Strip *in_strip = owner(lo);
                int hi = lo + 1;
                while (hi < xa + nr && owner(hi) == in_strip) hi++;
                if (in_strip) {
                    StepParams Q = P;
                    Q.src = in_strip->buf;
                    Q.plane = (long long)(in_strip->b - in_strip->a + 2) * c->pitch;
                    Q.sbase = (in_strip->a + c->NX - 1) % c->NX;
                    Q.out_cur = c->outbuf[2];
                    Q.use_snap = 0;
                    Q.row0a = lo;
                    Q.na = hi - lo;
                    e = launch<true, false, true, false>(Q, (hi - lo) * P.bpr, bs, c->stream);
                    c->launches++;
                } else {
                    e = deep_final(P, lo, hi - lo, c->last_depth);
                }
                lo = hi;
            }
            c->launches--;   // (counted once more below)
        } else if (c->last_depth > 1) {
            e = deep_final(P, xa, nr, c->last_depth);
            c->launches--;
        } else
            e = c->has_bc ? launch<true, false, true, false>(P, blocks, bs, c->stream) : launch<false, false, true, false>(P, blocks, bs, c->stream);
        if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "materialize kernel launch failed: %s", cudaGetErrorString(e));
        c->launches++;
        const size_t n = (size_t)nr * w, o = (size_t)(xa - x0) * w;
        if (dev_rho || dev_u) continue;   // device destination: nothing to copy, nothing to wait for
        if (to_host) {
            if (f) CK(cudaMemcpyAsync(f + o * 9, c->stage_f, n * 72, cudaMemcpyDeviceToHost, c->stream));
            if (rho) CK(cudaMemcpyAsync(rho + o, c->stage_rho, n * 8, cudaMemcpyDeviceToHost, c->stream));
            if (u) CK(cudaMemcpyAsync(u + o * 2, c->stage_u, n * 16, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        } else {
            const int nb = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
            k_minmax<<<nb, 256, 0, c->stream>>>((long long)n, c->stage_rho, c->stage_u, c->mm_acc);
            CK(cudaGetLastError());
            c->launches++;
            CK(cudaStreamSynchronize(c->stream));
        }
    }
    if (dev_rho || dev_u) return LBM_OK;
    return async_error(c, "lbm_materialize");
}

static int check_region(lbm_ctx *c, int x0, int x1, int y0, int y1, const char *who)
{
    if (!c) return fail(LBM_ERR_ARG, "%s: null context", who);
    if (!c->loaded || c->t == 0) return fail(LBM_ERR_STATE, "%s: no step taken since the state was loaded — the caller still holds it", who);
    if (x0 < 0 || y0 < 0 || x1 > c->NX || y1 > c->NY || x0 >= x1 || y0 >= y1) return fail(LBM_ERR_ARG, "%s: empty or out-of-range region", who);
    if (c->gx >= 2 && (x0 < c->gx || x1 > c->NX - c->gx)) return fail(LBM_ERR_ARG, "%s: slabs with %d ghost rows expose their interior rows [%d, nx-%d) only", who, c->gx, c->gx, c->gx);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream_edge));
    return LBM_OK;
}

extern "C" int lbm_materialize_region(lbm_ctx *c, int x0, int x1, int y0, int y1, double *f, double *rho, double *u)
{
    if (int rc = check_region(c, x0, x1, y0, y1, "lbm_materialize")) return rc;
    if (!f && !rho && !u) return LBM_OK;
    return materialize_rows(c, x0, x1, y0, y1, f, rho, u, true);
}

extern "C" int lbm_materialize(lbm_ctx *c, double *f, double *rho, double *u)
{
    if (!c) return fail(LBM_ERR_ARG, "lbm_materialize: null context");
    return lbm_materialize_region(c, 0, c->NX, 0, c->NY, f, rho, u);
}

// Publishes a halo epoch on its own (after the upload phase of a streamed run: every ghost store of the first collisions
// is complete when this kernel starts — stream order — and visible to the peers after the system-scope fence).
__global__ void k_halo_publish(const __grid_constant__ StepParams P)
{
    if (threadIdx.x == 0) {
        __threadfence_system();
        for (int s = 0; s < 9; s++)
            if (P.flag_out[s]) *(volatile unsigned *)P.flag_out[s] = P.signal_value;
        __threadfence_system();
    }
}

// ---- whole job from and to host memory, streamed ---------------------------------------------------------------
// lbm_run_host = lbm_upload + lbm_step(n) + lbm_materialize, same bits, same final state of the context — but on a fluid
// lattice without ghosts the three phases are PIPELINED over row chunks (time-skewed passes): as soon as rows [0, X) of
// the initial state are on the device, pass p can be completed on rows [pD, X - pD) (D = steps of a full pass), and
// rows [PD, X - PD) of the result can already travel back. With 256 MB chunks the host->device
// copy of chunk c+1, the passes over chunk c (a few hundred microseconds: hidden) and the device->host copy of the
// rows chunk c completed run concurrently on three streams; the job then takes about max(upload, download) on a
// full-duplex PCIe link instead of upload + compute + download. Level p lives in S[p % 2]: pass p on rows [a, b)
// overwrites level p-2 there, which pass p-1 has finished reading because it already completed rows up to b + d_p.
// The rows within s_p of the periodic seam (x = 0 / NX) are done last, after the last chunk has arrived.
static int streamed_setup(lbm_ctx *c, int n_levels)
{
    if (!c->sr) {
        c->sr = new lbm_ctx::Streamed;
        lbm_ctx::Streamed &R = *c->sr;
        R.in_f[0] = c->stage_f;
        R.in_rho[0] = c->stage_rho;
        R.in_u[0] = c->stage_u;
        CK(cudaMalloc(&R.in_f[1], c->stage_cells * 72));
        CK(cudaMalloc(&R.in_rho[1], c->stage_cells * 8));
        CK(cudaMalloc(&R.in_u[1], c->stage_cells * 16));
        for (int k = 0; k < 2; k++) {
            CK(cudaMalloc(&R.out_f[k], c->stage_cells * 72));
            CK(cudaMalloc(&R.out_rho[k], c->stage_cells * 8));
            CK(cudaMalloc(&R.out_u[k], c->stage_cells * 16));
            for (cudaEvent_t *e : {&R.in_ready[k], &R.in_free[k], &R.out_ready[k], &R.out_free[k]})
                CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        }
        CK(cudaStreamCreateWithFlags(&R.h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&R.d2h, cudaStreamNonBlocking));
    }
    if (c->sr->tclock_cap < n_levels) {
        if (c->sr->tclock) CK(cudaFree(c->sr->tclock));
        c->sr->tclock = nullptr;
        CK(cudaMalloc(&c->sr->tclock, (size_t)n_levels * 8));
        c->sr->tclock_cap = n_levels;
    }
    return LBM_OK;
}

extern "C" int lbm_run_host(lbm_ctx *c, const double *f, const double *rho, const double *u, double omega, int n_steps, double *f_out,
                            double *rho_out, double *u_out)
{
    if (!c || !f || !rho || !u) return fail(LBM_ERR_ARG, "lbm_run_host: null pointer");
    if (n_steps < 1) return fail(LBM_ERR_ARG, "lbm_run_host: n_steps < 1");
    const int D = fused_ok(c) ? max_depth(c) : 1;
    long long chunk_rows = c->stage_cells / c->NY;
    if (c->streamed_chunk_rows > 0) chunk_rows = std::min<long long>(chunk_rows, c->streamed_chunk_rows);   // (tests)
    // fluid lattices: one block with in-kernel periodic wrap, or a slab with >= D ghost rows whose neighbours are connected
    const int g = c->gx;
    const bool streamed = c->use_streamed && D >= 2 && !c->has_bc && !c->gy && (g == 0 ? !c->halo_ready : (g >= D && c->halo_ready)) &&
                          (long long)c->NX >= 2LL * (n_steps + D) + 2 * chunk_rows + 4LL * g && chunk_rows >= 8;
    if (!streamed) {   // every other lattice: the three calls
        if (c->any_remote)
            return fail(LBM_ERR_STATE, "lbm_run_host: this lattice has remote halo neighbours and cannot take the pipelined schedule "
                                       "(needs a fluid slab with >= %d ghost rows and nx >= %lld): use lbm_upload, a process-group barrier, "
                                       "lbm_step and lbm_materialize_region", D, 2LL * (n_steps + D) + 2 * chunk_rows + 4LL * g);
        if (int rc = lbm_upload(c, f, rho, u, omega)) return rc;
        if (int rc = lbm_step(c, omega, n_steps)) return rc;
        if (!f_out && !rho_out && !u_out) return LBM_OK;
        if (g >= 2) return lbm_materialize_region(c, g, c->NX - g, 0, c->NY, f_out ? f_out + (size_t)g * c->NY * 9 : nullptr,
                                                  rho_out ? rho_out + (size_t)g * c->NY : nullptr, u_out ? u_out + (size_t)g * c->NY * 2 : nullptr);
        return lbm_materialize(c, f_out, rho_out, u_out);
    }
    // pass plan: the schedule lbm_step would take
    std::vector<int> depth, s(1, 0);
    for (int left = n_steps; left > 0;) {
        const int d = left >= D ? D : left;   // remainder: one two-step pass or one one-step launch
        depth.push_back(d);
        s.push_back(s.back() + d);
        left -= d;
    }
    const int NP = (int)depth.size(), NX = c->NX, NY = c->NY;
    // Row lag of level p behind the upload front: p * D rows, also where a pass is shallower than D. (With the true
    // s_p a two-step pass after three-step passes would overwrite, in S[p % 2], one row of level p-2 that pass p-1 of
    // the NEXT chunk still has to read: the lag per pass must cover the deepest read halo.)
    std::vector<int> lag(NP + 1);
    for (int p = 0; p <= NP; p++) lag[p] = p * D;
    const int sP = lag[NP];
    InitParams Q;
    if (int rc = begin_load(c, omega, Q)) return rc;
    if (int rc = streamed_setup(c, NP + 1)) return rc;
    lbm_ctx::Streamed &R = *c->sr;
    const bool probe = c->probe && c->px >= 0;
    {
        std::vector<long long> tl(s.begin(), s.end());
        CK(cudaMemcpyAsync(R.tclock, tl.data(), tl.size() * 8, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync(c->tcount, 0, 32, c->stream));
        CK(cudaStreamSynchronize(c->stream));   // (tl is a host temporary)
        *c->progress = 0;
    }
    const bool want_out = f_out || rho_out || u_out;
    int out_k = 0, n_out = 0;
    const int I0 = g, I1 = NX - g;            // the rows this rank computes
    const bool remote = g && c->any_remote;

    // pass p (1-based) over lattice rows [lo, hi) and, optionally, a second range [lo2, hi2)
    auto run_pass = [&](int p, int lo, int hi, int lo2, int hi2) -> int {
        if (hi <= lo && hi2 <= lo2) return LBM_OK;
        if (hi <= lo) { lo = lo2; hi = hi2; lo2 = hi2 = 0; }
        StepParams P;
        fill_common(c, P, (p - 1) & 1, p & 1, omega);
        if (probe) {
            P.px = c->px;
            P.py = c->py;
            P.probe = c->probe;
            P.progress = c->progress;
            P.probe_cap = c->probe_cap;
            P.tc_in = R.tclock + (p - 1);
            P.tc_out = R.tclock + p;
        }
        const int d = depth[p - 1];
        if (d >= 2) return fused_launch<false>(c, P, lo, hi - lo, lo2, hi2 - lo2, 32, c->stream, d);
        if (int rc = rows_launch(c, P, lo, hi - lo, 0, 0, false, false, c->stream)) return rc;
        return hi2 > lo2 ? rows_launch(c, P, lo2, hi2 - lo2, 0, 0, false, false, c->stream) : LBM_OK;
    };
    // rows [lo, hi) of the result: the last pass again with its last level in FINAL mode -> staging -> host
    auto emit_rows = [&](int lo, int hi) -> int {
        if (!want_out || hi <= lo) return LBM_OK;
        const int k = out_k;
        out_k ^= 1;
        if (n_out >= 2) CK(cudaStreamWaitEvent(c->stream, R.out_free[k], 0));
        n_out++;
        StepParams P;
        fill_common(c, P, (NP - 1) & 1, NP & 1, omega);
        P.row0a = lo;
        P.na = hi - lo;
        P.y0 = 0;
        P.y1 = NY;
        P.ox0 = lo;
        P.oy0 = 0;
        P.ow = NY;
        P.o_f = f_out ? R.out_f[k] : nullptr;
        P.o_rho = rho_out ? R.out_rho[k] : nullptr;
        P.o_u = u_out ? R.out_u[k] : nullptr;
        const int d = depth[NP - 1];
        cudaError_t e;
        if (d >= 2) {
            const int W = deep_width(d);
            P.seg = std::min(hi - lo, 32);
            P.strip0 = 0;
            dim3 grid((NY + W - 1) / W, (hi - lo + P.seg - 1) / P.seg);
            deep_kernel(d, false, false, true)<<<grid, kDeepThreads, deep_smem(d), c->stream>>>(P);
            e = cudaGetLastError();
        } else {
            const int bs = block_size(NY);
            P.bpr = (NY + bs - 1) / bs;
            e = launch<false, false, true, false>(P, (hi - lo) * P.bpr, bs, c->stream);
        }
        if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "lbm_run_host: materialize launch failed: %s", cudaGetErrorString(e));
        c->launches++;
        CK(cudaEventRecord(R.out_ready[k], c->stream));
        CK(cudaStreamWaitEvent(R.d2h, R.out_ready[k], 0));
        const size_t n = (size_t)(hi - lo) * NY, o = (size_t)lo * NY;
        if (f_out) CK(cudaMemcpyAsync(f_out + o * 9, R.out_f[k], n * 72, cudaMemcpyDeviceToHost, R.d2h));
        if (rho_out) CK(cudaMemcpyAsync(rho_out + o, R.out_rho[k], n * 8, cudaMemcpyDeviceToHost, R.d2h));
        if (u_out) CK(cudaMemcpyAsync(u_out + o * 2, R.out_u[k], n * 16, cudaMemcpyDeviceToHost, R.d2h));
        CK(cudaEventRecord(R.out_free[k], R.d2h));
        return LBM_OK;
    };

    int chunk = 0;
    for (int X0 = 0; X0 < NX; X0 += (int)chunk_rows, chunk++) {
        const int X1 = (int)std::min<long long>(NX, X0 + chunk_rows), k = chunk & 1;
        const size_t n = (size_t)(X1 - X0) * NY, o = (size_t)X0 * NY;
        if (chunk >= 2) CK(cudaStreamWaitEvent(R.h2d, R.in_free[k], 0));
        CK(cudaMemcpyAsync(R.in_f[k], f + o * 9, n * 72, cudaMemcpyHostToDevice, R.h2d));
        CK(cudaMemcpyAsync(R.in_rho[k], rho + o, n * 8, cudaMemcpyHostToDevice, R.h2d));
        CK(cudaMemcpyAsync(R.in_u[k], u + o * 2, n * 16, cudaMemcpyHostToDevice, R.h2d));
        CK(cudaEventRecord(R.in_ready[k], R.h2d));
        CK(cudaStreamWaitEvent(c->stream, R.in_ready[k], 0));
        Q.in_f = R.in_f[k];
        Q.in_rho = R.in_rho[k];
        Q.in_u = R.in_u[k];
        if (int rc = first_collide(c, Q, X0, X1 - X0)) return rc;
        CK(cudaEventRecord(R.in_free[k], c->stream));
        // level p is now computable on rows [I0 + lag_p, min(X1, I1) - lag_p); new with this chunk: from X0 - lag_p on
        for (int p = 1; p <= NP; p++)
            if (int rc = run_pass(p, std::max(X0 - lag[p], I0 + lag[p]), std::min(X1, I1) - lag[p], 0, 0)) return rc;
        if (int rc = emit_rows(std::max(X0 - sP, I0 + sP), std::min(X1, I1) - sP)) return rc;
    }
    if (!g) {
        // the periodic seam: rows [NX - lag_p, NX) and [0, lag_p) of level p, in pass order
        for (int p = 1; p <= NP; p++)
            if (int rc = run_pass(p, NX - lag[p], NX, 0, lag[p])) return rc;
    } else {
        // Slab: the rows within lag_p of the slab edges. Everything so far touched neither my ghost rows nor the
        // neighbours' — except the first collisions, which stored the edge rows of S_0 into the neighbours' ghost rows:
        // publish that (one epoch), then take the passes in lockstep with the neighbours: the g rows next to each edge by
        // the HALO launch (waits for the neighbours' previous pass, stores into their ghost rows, publishes), the rest plain.
        if (remote) {
            StepParams Pp;
            fill_common(c, Pp, 1, 0, omega);
            fill_halo(c, Pp, 0, true);
            Pp.signal_value = c->halo_epoch + 1;
            k_halo_publish<<<1, 32, 0, c->stream>>>(Pp);
            CK(cudaGetLastError());
            c->halo_epoch++;
        }
        for (int p = 1; p <= NP; p++) {
            const int d = depth[p - 1];
            StepParams Pe;
            fill_common(c, Pe, (p - 1) & 1, p & 1, omega);
            fill_halo(c, Pe, p & 1, true);
            if (probe) {
                Pe.px = c->px;
                Pe.py = c->py;
                Pe.probe = c->probe;
                Pe.progress = c->progress;
                Pe.probe_cap = c->probe_cap;
                Pe.tc_in = R.tclock + (p - 1);
                Pe.tc_out = R.tclock + p;
            }
            if (remote) {
                Pe.wait_value = c->halo_epoch;
                Pe.signal_value = c->halo_epoch + 1;
            }
            if (d >= 2) {
                if (int rc = fused_launch<true>(c, Pe, I0, g, I1 - g, g, g, c->stream, d)) return rc;
            } else {
                if (int rc = rows_launch(c, Pe, I0, g, I1 - g, g, false, true, c->stream)) return rc;
            }
            if (remote) c->halo_epoch++;
            if (int rc = run_pass(p, I0 + g, I0 + lag[p], I1 - lag[p], I1 - g)) return rc;
        }
    }
    if (int rc = emit_rows(I1 - sP, I1)) return rc;
    if (int rc = emit_rows(I0, I0 + sP)) return rc;
    // the context is now where lbm_upload + lbm_step(n_steps) would have left it
    c->cur = NP & 1;
    c->t = n_steps;
    c->last_depth = depth[NP - 1];
    c->omega = omega;
    c->loaded = true;
    {
        long long tt[4] = {0, 0, 0, 0};
        tt[c->cur] = n_steps;
        tt[c->cur ^ 1] = n_steps - depth[NP - 1];
        tt[2] = tt[3] = n_steps;
        CK(cudaMemcpyAsync(c->tcount, tt, 32, cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(R.d2h));
    CK(cudaStreamSynchronize(R.h2d));
    return async_error(c, "lbm_run_host");
}

// ---- device-side history -------------------------------------------------------------------------------------
// Drivers that KEEP a field of every step (velocities.append(velocity), src/experiments.py:254, :542) and look at a few of
// them after the loop: the fields of a step are parked in a device slot by one asynchronous launch instead of travelling
// to the host behind a synchronisation after every step; a slot is read back only if somebody looks at it.
extern "C" int lbm_history_config(lbm_ctx *c, int n_slots)
{
    if (!c || n_slots < 0) return fail(LBM_ERR_ARG, "lbm_history_config: bad argument");
    if (c->gx >= 2) return fail(LBM_ERR_ARG, "lbm_history_config: slabs with ghost rows of the multi-step kernel keep no history");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->hist) CK(cudaFree(c->hist));
    c->hist = nullptr;
    c->n_hist = 0;
    if (n_slots == 0) return LBM_OK;
    const size_t bytes = (size_t)n_slots * 3 * c->NX * c->NY * 8;
    if (cudaMalloc(&c->hist, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(LBM_ERR_NOMEM, "cannot allocate %.1f MB for %d history slots", bytes / 1e6, n_slots);
    }
    c->n_hist = n_slots;
    return LBM_OK;
}

extern "C" int lbm_history_store(lbm_ctx *c, int slot)
{
    if (int rc = check_region(c, 0, c ? c->NX : 0, 0, c ? c->NY : 0, "lbm_history_store")) return rc;
    if (slot < 0 || slot >= c->n_hist) return fail(LBM_ERR_ARG, "lbm_history_store: slot %d of %d", slot, c->n_hist);
    double *base = c->hist + (size_t)slot * 3 * c->NX * c->NY;
    return materialize_rows(c, 0, c->NX, 0, c->NY, nullptr, nullptr, nullptr, false, base, base + (size_t)c->NX * c->NY);
}

extern "C" int lbm_history_read(lbm_ctx *c, int slot, double *rho, double *u)
{
    if (!c || slot < 0 || slot >= c->n_hist) return fail(LBM_ERR_ARG, "lbm_history_read: bad slot");
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->NX * c->NY;
    const double *base = c->hist + (size_t)slot * 3 * n;
    if (rho) CK(cudaMemcpyAsync(rho, base, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (u) CK(cudaMemcpyAsync(u, base + n, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return async_error(c, "lbm_history_read");
}

extern "C" int lbm_minmax(lbm_ctx *c, int x0, int x1, int y0, int y1, double out[4])
{
    if (int rc = check_region(c, x0, x1, y0, y1, "lbm_minmax")) return rc;
    if (!out) return fail(LBM_ERR_ARG, "lbm_minmax: out is null");
    const long long init[4] = {0x7fffffffffffffffLL, -0x7fffffffffffffffLL - 1, 0x7fffffffffffffffLL, -0x7fffffffffffffffLL - 1};
    CK(cudaMemcpyAsync(c->mm_acc, init, 32, cudaMemcpyHostToDevice, c->stream));
    if (int rc = materialize_rows(c, x0, x1, y0, y1, nullptr, nullptr, nullptr, false)) return rc;
    long long acc[4];
    CK(cudaMemcpy(acc, c->mm_acc, 32, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; i++) out[i] = unord(acc[i]);
    return LBM_OK;
}

extern "C" int lbm_probe_config(lbm_ctx *c, int x, int y, int capacity)
{
    if (!c) return fail(LBM_ERR_ARG, "lbm_probe_config: null context");
    if (x < 0 || y < 0 || x >= c->NX || y >= c->NY || capacity < 1) return fail(LBM_ERR_ARG, "lbm_probe_config: probe outside the lattice or capacity < 1");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream_edge));
    CK(cudaStreamSynchronize(c->stream));
    if (c->probe) CK(cudaFreeHost(c->probe));
    c->probe = nullptr;
    CK(cudaHostAlloc((void **)&c->probe, (size_t)capacity * 16, cudaHostAllocMapped));
    memset(c->probe, 0, (size_t)capacity * 16);
    *c->progress = c->t;
    c->px = x;
    c->py = y;
    c->probe_cap = capacity;
    // S[cur] holds time t, S[cur^1] time t-1 (the launch an omega change redoes reads it and must land on t again)
    long long tt[4] = {c->t, c->t, c->t, c->t};
    tt[c->cur ^ 1] = c->t - 1;
    CK(cudaMemcpyAsync(c->tcount, tt, 32, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    drop_graphs(c);
    return LBM_OK;
}

extern "C" int lbm_probe_read(lbm_ctx *c, int64_t t0, int n, double *uxuy)
{
    if (!c || !uxuy) return fail(LBM_ERR_ARG, "lbm_probe_read: null pointer");
    if (!c->probe) return fail(LBM_ERR_STATE, "lbm_probe_read: no probe configured");
    if (n < 0 || t0 < 1 || t0 + n - 1 > c->t || c->t - t0 >= c->probe_cap)
        return fail(LBM_ERR_ARG, "lbm_probe_read: steps [%lld, %lld) not in the ring (time %lld, capacity %d)", (long long)t0,
                    (long long)t0 + n, c->t, c->probe_cap);
    CK(cudaSetDevice(c->device));
    // Wait for the newest requested sample only: the device may be many steps further down its queue. (When both
    // streams have drained, everything that was ever going to be recorded is in the ring.)
    const long long target = t0 + n - 1;
    volatile long long *progress = c->progress;
    while (*progress < target && !*(volatile unsigned *)c->err_host) {
        const cudaError_t a = cudaStreamQuery(c->stream_edge), b = cudaStreamQuery(c->stream);
        if (a != cudaSuccess && a != cudaErrorNotReady) CK(a);
        if (b != cudaSuccess && b != cudaErrorNotReady) CK(b);
        if (a == cudaSuccess && b == cudaSuccess) break;
    }
    __sync_synchronize();
    if (int rc = async_error(c, "lbm_probe_read")) return rc;
    for (int i = 0; i < n;) {
        const int slot = (int)((t0 + i) % c->probe_cap);
        const int run = std::min(n - i, c->probe_cap - slot);
        memcpy(uxuy + 2 * i, c->probe + 2 * slot, (size_t)run * 16);
        i += run;
    }
    return LBM_OK;
}

// ---- halo ----------------------------------------------------------------------------------------------
extern "C" int lbm_halo_export_handle(lbm_ctx *c, lbm_halo_export *out)
{
    if (!c || !out) return fail(LBM_ERR_ARG, "lbm_halo_export_handle: null pointer");
    memset(out, 0, sizeof *out);
    CK(cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) <= LBM_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, c->arena) == cudaSuccess)
        memcpy(out->mem_handle, &h, sizeof h);
    else
        cudaGetLastError();   // IPC unavailable (e.g. some containers): same-process neighbours still work
    out->device = c->device;
    out->nx = c->NX;
    out->ny = c->NY;
    out->pitch = c->pitch;
    out->pid = (int64_t)getpid();
    out->arena_ptr = (uint64_t)(uintptr_t)c->arena;
    out->arena_bytes = (int64_t)c->arena_bytes;
    return LBM_OK;
}

extern "C" int lbm_halo_connect(lbm_ctx *c, int slot, const lbm_halo_export *peer)
{
    if (!c || !peer) return fail(LBM_ERR_ARG, "lbm_halo_connect: null pointer");
    if (slot < 0 || slot > 8 || slot == 4) return fail(LBM_ERR_ARG, "lbm_halo_connect: slot must be 0..8 except 4");
    const int dx = slot / 3 - 1, dy = slot % 3 - 1;
    if ((dx && !c->gx) || (dy && !c->gy)) return fail(LBM_ERR_ARG, "lbm_halo_connect: no ghost layer in that direction");
    // an x-neighbour shares my y extent and vice versa (Cartesian blocks, parallelization_utils.py:111-120)
    if (dx == 0 && peer->nx != c->NX) return fail(LBM_ERR_ARG, "lbm_halo_connect: y-neighbour must have the same nx");
    if (dy == 0 && peer->ny != c->NY) return fail(LBM_ERR_ARG, "lbm_halo_connect: x-neighbour must have the same ny");
    CK(cudaSetDevice(c->device));
    Peer &pr = c->peer[slot];
    pr.nx = peer->nx;
    pr.ny = peer->ny;
    pr.pitch = peer->pitch;
    if (peer->pid == (int64_t)getpid()) {
        pr.arena = (char *)(uintptr_t)peer->arena_ptr;
        pr.remote = pr.arena != c->arena;
        if (pr.remote && peer->device != c->device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, c->device, peer->device));
            if (!can) return fail(LBM_ERR_CUDA, "device %d cannot access device %d", c->device, peer->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
        }
    } else {
        // reuse a mapping of the same peer opened for another slot (cudaIpcOpenMemHandle is once per process)
        if (pr.mapped) ipc_release(pr.mapped);   // slot re-connected
        pr.mapped = nullptr;
        void *m = nullptr;
        if (int rc = ipc_acquire(peer->mem_handle, &m)) return rc;
        pr.mapped = m;
        pr.arena = (char *)m;
        pr.remote = true;
    }
    pr.connected = true;
    return LBM_OK;
}

extern "C" int lbm_halo_finalize(lbm_ctx *c)
{
    if (!c) return fail(LBM_ERR_ARG, "lbm_halo_finalize: null context");
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++) {
            if (!dx && !dy) continue;
            const bool needed = (!dx || c->gx) && (!dy || c->gy);
            if (needed && !c->peer[(dx + 1) * 3 + dy + 1].connected)
                return fail(LBM_ERR_STATE, "lbm_halo_finalize: neighbour (%d,%d) not connected", dx, dy);
        }
    c->any_remote = false;
    for (int s = 0; s < 9; s++) c->any_remote |= c->peer[s].connected && c->peer[s].remote;
    c->halo_ready = true;
    c->halo_epoch = 0;
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->arena + c->off_flags, 0, 256, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return LBM_OK;
}
