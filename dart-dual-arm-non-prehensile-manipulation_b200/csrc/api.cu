// C ABI (include/dart_b200.h) over the sm_100a kernels.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <string.h>
#include <stdlib.h>
#include <new>

#include "launch.h"
#include "opts.h"

using namespace dart;

struct dart_solver {
    dart_cfg cfg;
    int device;
    SolverOpts opts;
    int64_t launches;
    LaunchInfo last;
    // staging for the *_host entry point
    cudaStream_t stream;
    void* pin;   size_t pin_bytes;
    void* dev;   size_t dev_bytes;
    double* rows;      // dart_set_result_rows
    int32_t rows_cap;
    double* dual;      // dart_set_dual_state
    int32_t dual_cap;
    // hand-over area of the two-axis models (KArgs::axis_part / axis_sync), grown on demand
    double* axis_part;
    int32_t* axis_sync;
    int32_t axis_cap;
    double* peer_rows[DART_MAX_PEERS];     // dart_set_result_rows_peers
    int32_t n_peers;
    int64_t peer_off;
    int32_t barrier;   // dart_set_barrier_strategy: -1 = default of fill_opts
};

extern "C" int dart_default_cfg(int32_t method, dart_cfg* c) {
    if (!c) return DART_ERR_ARG;
    memset(c, 0, sizeof(*c));
    c->method = method;
    c->Ts = 0.002;
    c->tol = 1e-8;
    c->mu_init = 0.0;          // the barrier strategy's default: 0.1 monotone (IPOPT's mu_init), 0.01 predictor-corrector
    c->max_iter = 200;
    switch (method) {
        case DART_PMPC:   // PMPC/main.py:59-69
            c->N = 15; c->g = -9.81; c->u_lo = -0.6; c->u_hi = 0.6; c->Qp = 400.0; c->Qv = 2.0; c->R = 0.2; c->mu = 0.1;
            return DART_OK;
        case DART_RMPC:   // RMPC/dev_dual/rob_ctrl.py:281-284
            c->N = 20; c->g = -9.81; c->u_lo = -0.6; c->u_hi = 0.6; c->du_lo = -0.06; c->du_hi = 0.06;
            c->vmax = 0.2; c->v_eps = 0.1; c->Qp = 80.0; c->Qv = 2.0; c->R = 0.02; c->Rdu = 1.0;
            return DART_OK;
        case DART_LMPC: { // LMPC/src/run.py:118-126
            c->N = 20; c->g = 9.81; c->u_lo = -0.4; c->u_hi = 0.4;
            const double Q[8] = {200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0};
            for (int i = 0; i < 8; ++i) { c->Q[i] = Q[i]; c->Qt[i] = Q[i]; }
            c->Rl[0] = 0.1; c->Rl[1] = 0.1; c->Rl[2] = 1.0; c->Rl[3] = 1.0;
            return DART_OK;
        }
        default: return DART_ERR_ARG;
    }
}

static int sizes(const dart_cfg& c, int& nx, int& nref, int& naux, int& nw) {
    switch (c.method) {
        case DART_PMPC: nx = 6; nref = 6; naux = 4; nw = PmpcAxis::nw(c.N); return DART_OK;
        case DART_RMPC: nx = 4; nref = (c.N + 1) * 4; naux = 16; nw = Rmpc::nw(c.N); return DART_OK;
        case DART_LMPC: nx = 8; nref = 8; naux = 36; nw = LmpcAxis::nw(c.N); return DART_OK;
        default: return DART_ERR_ARG;
    }
}

extern "C" int dart_create(dart_handle* out, const dart_cfg* cfg, int device) {
    if (!out || !cfg) return DART_ERR_ARG;
    int nx, nref, naux, nw;
    if (sizes(*cfg, nx, nref, naux, nw) != DART_OK) return DART_ERR_ARG;
    if (cfg->N < 1 || cfg->N > 64 || !(cfg->Ts > 0.0) || !(cfg->u_hi > cfg->u_lo)) return DART_ERR_ARG;
    if (cfg->method == DART_RMPC && (!(cfg->du_hi > cfg->du_lo) || !(cfg->vmax > 0.0) || !(cfg->v_eps > 0.0))) return DART_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return DART_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) return DART_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return DART_ERR_CUDA;
    dart_solver* h = new (std::nothrow) dart_solver();
    if (!h) return DART_ERR_ALLOC;
    h->cfg = *cfg;
    h->device = device;
    fill_opts(*cfg, h->opts);
    h->launches = 0;
    memset(&h->last, 0, sizeof(h->last));
    h->pin = nullptr; h->pin_bytes = 0; h->dev = nullptr; h->dev_bytes = 0; h->rows = nullptr; h->rows_cap = 0; h->dual = nullptr; h->dual_cap = 0;
    h->axis_part = nullptr; h->axis_sync = nullptr; h->axis_cap = 0;
    h->n_peers = 0; h->peer_off = 0;
    h->barrier = -1;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return DART_ERR_CUDA; }
    *out = h;
    return DART_OK;
}

extern "C" int dart_destroy(dart_handle h) {
    if (!h) return DART_ERR_ARG;
    cudaSetDevice(h->device);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->dev) cudaFree(h->dev);
    if (h->axis_part) cudaFree(h->axis_part);
    if (h->axis_sync) cudaFree(h->axis_sync);
    cudaStreamDestroy(h->stream);
    delete h;
    return DART_OK;
}

static int dual_doubles(const dart_cfg& c) {
    switch (c.method) {
        case DART_PMPC: return PmpcAxis::NAXIS * Solver<PmpcAxis, HostTile, 0>::dual_doubles(c.N);
        case DART_RMPC: return Rmpc::NAXIS * Solver<Rmpc, HostTile, 0>::dual_doubles(c.N);
        case DART_LMPC: return LmpcAxis::NAXIS * Solver<LmpcAxis, HostTile, 0>::dual_doubles(c.N);
    }
    return -1;
}

extern "C" int dart_ndual(dart_handle h) { return h ? dual_doubles(h->cfg) : DART_ERR_ARG; }

extern "C" int dart_set_dual_state(dart_handle h, double* dual, int32_t capacity_rows) {
    if (!h || (dual && capacity_rows <= 0)) return DART_ERR_ARG;
    h->dual = dual;
    h->dual_cap = dual ? capacity_rows : 0;
    return DART_OK;
}

extern "C" int dart_set_mu_init(dart_handle h, double mu_init) {
    if (!h || !(mu_init >= 0.0)) return DART_ERR_ARG;
    h->cfg.mu_init = mu_init;                  // 0 selects the strategy's default (0.1 / 0.01)
    fill_opts(h->cfg, h->opts);
    if (h->barrier >= 0) h->opts.mehrotra = h->barrier;
    return DART_OK;
}

extern "C" int dart_set_barrier_strategy(dart_handle h, int32_t strategy) {
    if (!h || (strategy != DART_BARRIER_MONOTONE && strategy != DART_BARRIER_MEHROTRA && strategy != DART_BARRIER_AUTO)) return DART_ERR_ARG;
    h->barrier = strategy;
    h->opts.mehrotra = strategy;
    return DART_OK;
}

extern "C" int dart_nx(dart_handle h) { int a, b, c, d; if (!h || sizes(h->cfg, a, b, c, d)) return DART_ERR_ARG; return a; }
extern "C" int dart_nref(dart_handle h) { int a, b, c, d; if (!h || sizes(h->cfg, a, b, c, d)) return DART_ERR_ARG; return b; }
extern "C" int dart_naux(dart_handle h) { int a, b, c, d; if (!h || sizes(h->cfg, a, b, c, d)) return DART_ERR_ARG; return c; }
extern "C" int dart_nw(dart_handle h) { int a, b, c, d; if (!h || sizes(h->cfg, a, b, c, d)) return DART_ERR_ARG; return d; }

extern "C" int dart_solve(dart_handle h, int32_t B, const double* x0, const double* ref, const double* aux,
                          const double* warm_w, double* w_out, double* u0_out, double* J_out, int32_t* status,
                          int32_t* iters, void* stream) {
    if (!h || B < 0 || !x0 || !ref || !u0_out || !J_out) return DART_ERR_ARG;
    if (h->cfg.method != DART_PMPC && !aux) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    if (h->rows && B > h->rows_cap) return DART_ERR_ARG;       // the registered result-row buffer is too small
    if (h->dual && B > h->dual_cap) return DART_ERR_ARG;       // the registered dual-state buffer is too small
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != h->device) return DART_ERR_ARG;   // launch from the handle's device
    cudaStream_t st = (cudaStream_t)stream;
    if (h->cfg.method == DART_LMPC && B > h->axis_cap) {
        // first solve of this size: (re)allocate the axes' hand-over area.  The counters are zero between launches
        // (the kernel re-arms them).  Allocation synchronises the device: size the handle with one warm-up solve
        // before capturing solves into a CUDA graph.
        if (cudaDeviceSynchronize() != cudaSuccess) return DART_ERR_CUDA;
        if (h->axis_part) cudaFree(h->axis_part);
        if (h->axis_sync) cudaFree(h->axis_sync);
        h->axis_part = nullptr; h->axis_sync = nullptr; h->axis_cap = 0;
        const size_t cap = (size_t)B + (size_t)B / 4 + 64;
        if (cudaMalloc(&h->axis_part, cap * LmpcAxis::NAXIS * 4 * sizeof(double)) != cudaSuccess) return DART_ERR_ALLOC;
        if (cudaMalloc(&h->axis_sync, cap * sizeof(int32_t)) != cudaSuccess) { cudaFree(h->axis_part); h->axis_part = nullptr; return DART_ERR_ALLOC; }
        if (cudaMemset(h->axis_sync, 0, cap * sizeof(int32_t)) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return DART_ERR_CUDA;
        h->axis_cap = (int32_t)cap;
    }
    KArgs a;
    a.B = B; a.N = h->cfg.N; a.o = h->opts; a.cfg = h->cfg;
    a.x0 = x0; a.ref = ref; a.aux = aux; a.warm = warm_w;
    launch_opts(h->cfg, warm_w != nullptr, a.o);
    a.w_out = w_out; a.u0 = u0_out; a.J = J_out; a.status = status; a.iters = iters; a.rows = h->rows; a.dual = h->dual;
    a.axis_part = h->axis_part; a.axis_sync = h->axis_sync;
    a.n_peers = h->n_peers; a.peer_off = h->peer_off;
    for (int p = 0; p < DART_MAX_PEERS; ++p) a.peer_rows[p] = p < h->n_peers ? h->peer_rows[p] : nullptr;
    int rc = launch_solve(a, h->cfg.lanes, h->cfg.block_threads, st, &h->last);
    if (rc != DART_OK) return rc;
    h->launches += 1;
    if (h->cfg.method == DART_PMPC && w_out) {
        rc = launch_pmpc_z(a, st);
        if (rc != DART_OK) return rc;
        h->launches += 1;
    }
    return DART_OK;
}

// DART_HOST_STAGED=1: always stage through device memory (A/B switch of the zero-copy path of dart_solve_host)
static bool getenv_staged() {
    static int v = -1;
    if (v < 0) v = getenv("DART_HOST_STAGED") ? 1 : 0;
    return v == 1;
}

static int ensure(dart_solver* h, size_t bytes) {
    if (bytes <= h->pin_bytes) return DART_OK;
    if (h->pin) cudaFreeHost(h->pin);
    if (h->dev) cudaFree(h->dev);
    h->pin = nullptr; h->dev = nullptr; h->pin_bytes = h->dev_bytes = 0;
    size_t cap = bytes + bytes / 4 + 4096;
    if (cudaMallocHost(&h->pin, cap) != cudaSuccess) return DART_ERR_ALLOC;
    if (cudaMalloc(&h->dev, cap) != cudaSuccess) { cudaFreeHost(h->pin); h->pin = nullptr; return DART_ERR_ALLOC; }
    h->pin_bytes = h->dev_bytes = cap;
    return DART_OK;
}

extern "C" int dart_solve_host(dart_handle h, int32_t B, const double* x0, const double* ref, const double* aux,
                               const double* warm_w, double* w_out, double* u0_out, double* J_out,
                               int32_t* status, int32_t* iters) {
    if (!h || B < 0 || !x0 || !ref || !u0_out || !J_out) return DART_ERR_ARG;
    if (h->cfg.method != DART_PMPC && !aux) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    if (cudaSetDevice(h->device) != cudaSuccess) return DART_ERR_CUDA;
    int nx, nref, naux, nw;
    sizes(h->cfg, nx, nref, naux, nw);
    // one packed input block [x0 | ref | aux | warm] and one packed output block [u0 | J | w | status | iters]
    const size_t n_x0 = (size_t)B * nx, n_ref = (size_t)B * nref, n_aux = aux ? (size_t)B * naux : 0,
                 n_warm = warm_w ? (size_t)B * nw : 0;
    const size_t in_d = n_x0 + n_ref + n_aux + n_warm;
    const size_t n_w = w_out ? (size_t)B * nw : 0;
    const size_t out_d = (size_t)B * 2 + B + n_w;
    const size_t out_i = (size_t)B * 2;
    const size_t bytes = (in_d + out_d) * sizeof(double) + out_i * sizeof(int32_t);
    int rc = ensure(h, bytes);
    if (rc != DART_OK) return rc;
    double* pin = (double*)h->pin;
    double* dev = (double*)h->dev;
    size_t o = 0;
    memcpy(pin + o, x0, n_x0 * 8); const size_t o_x0 = o; o += n_x0;
    memcpy(pin + o, ref, n_ref * 8); const size_t o_ref = o; o += n_ref;
    size_t o_aux = 0, o_warm = 0;
    if (aux) { memcpy(pin + o, aux, n_aux * 8); o_aux = o; o += n_aux; }
    if (warm_w) { memcpy(pin + o, warm_w, n_warm * 8); o_warm = o; o += n_warm; }
    // Small batches (the headline 1152-instance batch moves 147 kB in and 37 kB out): the kernel reads its inputs from and
    // writes its results to the PINNED staging block directly (mapped host memory under unified addressing) -- each input
    // row is read exactly once, at the start of its instance's solve, so the transfer is the same bytes over the same link
    // without two copy-engine launches and their serialisation in front of and behind a 0.08 ms kernel.  Larger batches, and
    // calls that carry plans (warm_w / w_out, re-read by the kernels), go through device staging with the copy engines.
    const bool zero_copy = !warm_w && !w_out && (in_d + out_d) * sizeof(double) <= (size_t)(1u << 20) && !getenv_staged();
    if (zero_copy) dev = pin;
    else if (cudaMemcpyAsync(dev, pin, in_d * 8, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return DART_ERR_CUDA;
    double* d_out = dev + in_d;
    double* d_u0 = d_out;
    double* d_J = d_out + (size_t)B * 2;
    double* d_w = w_out ? d_out + (size_t)B * 3 : nullptr;
    int32_t* d_st = (int32_t*)(d_out + out_d);
    int32_t* d_it = d_st + B;
    rc = dart_solve(h, B, dev + o_x0, dev + o_ref, aux ? dev + o_aux : nullptr, warm_w ? dev + o_warm : nullptr,
                    d_w, d_u0, d_J, d_st, d_it, (void*)h->stream);
    if (rc != DART_OK) return rc;
    double* p_out = pin + in_d;
    if (!zero_copy && cudaMemcpyAsync(p_out, d_out, out_d * 8 + out_i * 4, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return DART_ERR_CUDA;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return DART_ERR_CUDA;
    memcpy(u0_out, p_out, (size_t)B * 2 * 8);
    memcpy(J_out, p_out + (size_t)B * 2, (size_t)B * 8);
    if (w_out) memcpy(w_out, p_out + (size_t)B * 3, n_w * 8);
    const int32_t* p_i = (const int32_t*)(p_out + out_d);
    if (status) memcpy(status, p_i, (size_t)B * 4);
    if (iters) memcpy(iters, p_i + B, (size_t)B * 4);
    return DART_OK;
}

extern "C" int dart_pmpc_episode(dart_handle h, int32_t B, int32_t T, double* state, const double* target, const double* aux,
                                 const double* mu_plant, const double* coulomb, double tol, int32_t* nsteps,
                                 double* conv_time, double* effort, double* err, double* u0, double* J, int32_t* status,
                                 int32_t* iters, uint64_t* counters, void* stream) {
    if (!h || h->cfg.method != DART_PMPC || B < 0 || T < 0 || !state || !target || !mu_plant || !nsteps || !conv_time || !effort ||
        !err || !u0 || !J || !status || !iters || !counters)
        return DART_ERR_ARG;
    if (B == 0 || T == 0) return DART_OK;
    if (h->rows || h->dual) return DART_ERR_UNSUPPORTED;     // the episode kernel keeps no result rows / dual state
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != h->device) return DART_ERR_ARG;
    KArgs a;
    a.B = B; a.N = h->cfg.N; a.o = h->opts; a.cfg = h->cfg;
    a.x0 = state; a.ref = target; a.aux = aux; a.warm = nullptr;
    a.o.cold = 1;
    a.w_out = nullptr; a.u0 = u0; a.J = J; a.status = status; a.iters = iters; a.rows = nullptr; a.dual = nullptr; a.axis_part = nullptr; a.axis_sync = nullptr; a.n_peers = 0; a.peer_off = 0;
    PlantArgs pl{B, h->cfg.Ts, h->cfg.g, tol, mu_plant, coulomb, u0, target, state, conv_time, effort, err, nsteps};
    int rc = launch_episode_pmpc(a, T, pl, (unsigned long long*)counters, h->cfg.lanes, (cudaStream_t)stream, &h->last);
    if (rc != DART_OK) return rc;
    h->launches += 1;
    return DART_OK;
}

extern "C" int dart_set_result_rows(dart_handle h, double* rows, int32_t capacity_rows) {
    if (!h || (rows && capacity_rows <= 0)) return DART_ERR_ARG;
    h->rows = rows;
    h->rows_cap = rows ? capacity_rows : 0;
    return DART_OK;
}

extern "C" int dart_set_result_rows_peers(dart_handle h, double* const* peers, int32_t n_peers, int64_t row_offset) {
    if (!h || n_peers < 0 || n_peers > DART_MAX_PEERS || (n_peers > 0 && !peers) || row_offset < 0) return DART_ERR_ARG;
    for (int p = 0; p < n_peers; ++p) {
        if (!peers[p]) return DART_ERR_ARG;
        h->peer_rows[p] = peers[p];
    }
    h->n_peers = n_peers;
    h->peer_off = row_offset;
    return DART_OK;
}

extern "C" int dart_enable_peer_access(int device, int peer) {
    if (device == peer) return DART_OK;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, device, peer) != cudaSuccess || !can) { cudaGetLastError(); return DART_ERR_UNSUPPORTED; }
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) return DART_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return DART_ERR_CUDA;
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    cudaSetDevice(cur);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return DART_ERR_CUDA; }
    cudaGetLastError();
    return DART_OK;
}

// ---- peer-shared device buffers (CUDA IPC): dedicated cudaMalloc blocks, so the exported handle maps exactly the buffer
extern "C" int dart_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle64) {
    if (bytes <= 0 || !ptr || !handle64) return DART_ERR_ARG;
    void* p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return DART_ERR_ALLOC; }
    if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaFree(p); return DART_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return DART_ERR_UNSUPPORTED; }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return DART_OK;
}
// opened on the CURRENT device with lazy peer access: kernels of this device may then load / store the peer's buffer
extern "C" int dart_peer_open(const uint8_t* handle64, void** ptr) {
    if (!handle64 || !ptr) return DART_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return DART_ERR_UNSUPPORTED; }
    *ptr = p;
    return DART_OK;
}
extern "C" int dart_peer_close(void* ptr) {
    if (!ptr) return DART_ERR_ARG;
    return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}
extern "C" int dart_peer_free(void* ptr) {
    if (!ptr) return DART_ERR_ARG;
    return cudaFree(ptr) == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int64_t dart_launch_count(dart_handle h) { return h ? h->launches : -1; }

extern "C" int dart_last_launch_config(dart_handle h, int32_t* lanes, int32_t* block_threads, int32_t* grid, int32_t* smem_bytes) {
    if (!h) return DART_ERR_ARG;
    if (lanes) *lanes = h->last.lanes;
    if (block_threads) *block_threads = h->last.block_threads;
    if (grid) *grid = h->last.grid;
    if (smem_bytes) *smem_bytes = h->last.smem_bytes;
    return DART_OK;
}

extern "C" int dart_tilt_to_quat(int32_t B, const double* u, double* quat, void* stream) {
    if (B < 0 || !u || !quat) return DART_ERR_ARG;
    return launch_tilt_to_quat(B, u, quat, (cudaStream_t)stream);
}
