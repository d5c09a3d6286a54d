// TEST HARNESS ONLY -- compiles the device solver source (solver_core.cuh, models.cuh) for the host with a
// one-lane tile so that the interior-point/Riccati logic can be unit-tested on machines without a GPU.
// It is built into tests/hostemu/_build/ by tests/hostemu/build.py, is never loaded by the product package,
// and is not a CPU fallback: the C ABI in csrc/api.cu has no path to it.
#include <vector>
#include "../../dart-dual-arm-non-prehensile-manipulation_b200/csrc/models.cuh"
#include "../../dart-dual-arm-non-prehensile-manipulation_b200/csrc/opts.h"

using namespace dart;

template <class M>
static void run_all(const KArgs& a) {
    HostTile tile;
    std::vector<double> ws(Workspace<M>::doubles(a.N) + kSlot + 8);
    const BlockCtx bc{ws.data(), 0, 1, 0};          // a block of one problem: the same thread runs all three phases
    for (int inst = 0; inst < a.B; ++inst) {
        double Js = 0.0;
        int32_t st = 0, itx = 0;
        for (int ax = 0; ax < M::NAXIS; ++ax) {
            double J = 0, kkt = 0;
            int32_t s = 0, it = 0;
            // same dispatch as the device launcher: compile-time horizon when it is the reference's
            if (a.N == M::NDEF) solve_one<M, HostTile, M::NDEF>(tile, a, inst, ax, true, true, bc, ws.data() + kSlot, J, s, it, kkt);
            else solve_one<M, HostTile, 0>(tile, a, inst, ax, true, true, bc, ws.data() + kSlot, J, s, it, kkt);
            Js += J;
            st = s > st ? s : st;
            itx = it > itx ? it : itx;
        }
        a.J[inst] = Js;
        if (a.status) a.status[inst] = st;
        if (a.iters) a.iters[inst] = itx;
        if (a.cfg.method == DART_PMPC) pmpc_z_rollout(a, inst);
    }
}

static int solve_impl(const dart_cfg* cfg, int B, const double* x0, const double* ref, const double* aux,
                      const double* warm, double* w_out, double* u0, double* J, int32_t* status, int32_t* iters, double* dual);

extern "C" int hostemu_solve(const dart_cfg* cfg, int B, const double* x0, const double* ref, const double* aux,
                             const double* warm, double* w_out, double* u0, double* J, int32_t* status, int32_t* iters) {
    return solve_impl(cfg, B, x0, ref, aux, warm, w_out, u0, J, status, iters, nullptr);
}

// with a dual-state buffer [B, hostemu_ndual(cfg)] (dart_set_dual_state semantics)
extern "C" int hostemu_solve_dual(const dart_cfg* cfg, int B, const double* x0, const double* ref, const double* aux,
                                  const double* warm, double* w_out, double* u0, double* J, int32_t* status, int32_t* iters,
                                  double* dual) {
    return solve_impl(cfg, B, x0, ref, aux, warm, w_out, u0, J, status, iters, dual);
}

extern "C" int hostemu_ndual(const dart_cfg* cfg) {
    switch (cfg->method) {
        case DART_PMPC: return PmpcAxis::NAXIS * Solver<PmpcAxis, HostTile, 0>::dual_doubles(cfg->N);
        case DART_RMPC: return Rmpc::NAXIS * Solver<Rmpc, HostTile, 0>::dual_doubles(cfg->N);
        case DART_LMPC: return LmpcAxis::NAXIS * Solver<LmpcAxis, HostTile, 0>::dual_doubles(cfg->N);
    }
    return -1;
}

static int solve_impl(const dart_cfg* cfg, int B, const double* x0, const double* ref, const double* aux,
                      const double* warm, double* w_out, double* u0, double* J, int32_t* status, int32_t* iters, double* dual) {
    KArgs a;
    a.B = B; a.N = cfg->N; a.cfg = *cfg;
    fill_opts(*cfg, a.o);
    launch_opts(*cfg, warm != nullptr, a.o);
    a.x0 = x0; a.ref = ref; a.aux = aux; a.warm = warm; a.w_out = w_out; a.u0 = u0; a.J = J; a.status = status; a.iters = iters; a.rows = nullptr; a.dual = dual; a.axis_part = nullptr; a.axis_sync = nullptr; a.n_peers = 0; a.peer_off = 0;
    switch (cfg->method) {
        case DART_PMPC: run_all<PmpcAxis>(a); return 0;
        case DART_RMPC: run_all<Rmpc>(a); return 0;
        case DART_LMPC: run_all<LmpcAxis>(a); return 0;
    }
    return -1;
}
