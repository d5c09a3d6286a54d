// LMPC surrogate plant: one RK4 step, tilt held over Ts, of the controller's own 8-state model
// (LMPC/src/controller/rlmpc2.py:260-436) evaluated with a per-instance "true" 34-parameter vector that the controller
// does not know -- the thing the parameter-adaptation policy is trained to track.  Same device functions as the solver
// (LmpcAxis::load / rk4_sens), so plant and model agree to the last bit when the parameter vectors are equal.
#include <cuda_runtime.h>

#include "models.cuh"

namespace {
using dart::KArgs;
using dart::LmpcAxis;

__global__ void __launch_bounds__(128) lmpc_plant_step_kernel(const KArgs a, const double* __restrict__ u,
                                                              double* __restrict__ state_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
#pragma unroll
    for (int axis = 0; axis < 2; ++axis) {
        LmpcAxis::Prm p;
        LmpcAxis::load(p, a, b, axis);
        double x[4], F[4], A[16], Bm[4], tanu[1];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = a.x0[(size_t)b * 8 + LmpcAxis::xmap(axis, i)];
        const double ua = u[(size_t)b * 2 + axis];
        dart::rk4_sens<LmpcAxis>(p, x, &ua, p.Ts, F, A, Bm, tanu);
#pragma unroll
        for (int i = 0; i < 4; ++i) state_out[(size_t)b * 8 + LmpcAxis::xmap(axis, i)] = F[i];
    }
}

// End of one LMPCBatch closed-loop step, fused (what was ~10 eager tensor ops per step): the facade's "no fresh solution"
// branch (rlmpc2.py:1013-1018: next entry of the last good plan, which stays the warm start; last_control when there is no
// plan yet) and the hand-over of the command to the next step (u_prev and the u_prev columns of the solver's aux rows).
// One warp per instance; its lanes move the plan / the warm start.
struct PostArgs {
    int B, N, nw;
    const int32_t* status;
    const unsigned char* fresh;
    const double* w_prev;
    double *w_next, *u0, *u_prev, *aux, *plan_U;
    long long* plan_pos;
    unsigned char* have_plan;
    unsigned long long* n_fallback;
};

__global__ void __launch_bounds__(128) lmpc_post_step_kernel(const PostArgs a) {
    const int b = (int)((blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (b >= a.B) return;
    double* u0 = a.u0 + (size_t)b * 2;
    if (a.plan_U != nullptr) {
        const int32_t st = a.status[b];
        const bool had = a.have_plan[b] != 0;
        bool ok = (st == dart::ST_CONVERGED) || (st == dart::ST_ACCEPTABLE) || !had;   // no plan yet: take the iterate as it is
        if (a.fresh != nullptr) ok = ok && a.fresh[b] != 0;
        double* plan = a.plan_U + (size_t)b * a.N * 2;
        double* wn = a.w_next + (size_t)b * a.nw;
        long long pos = a.plan_pos[b];
        if (ok) {
            const double* U_new = wn + (size_t)(a.N + 1) * 8;
            for (int i = lane; i < a.N * 2; i += 32) plan[i] = U_new[i];
            pos = 0;                                  // the command is the solver's own u0 (= U_new[0])
        } else {
            const double* wp = a.w_prev + (size_t)b * a.nw;
            for (int i = lane; i < a.nw; i += 32) wn[i] = wp[i];      // the old plan stays the warm start
            pos = pos + 1 < a.N - 1 ? pos + 1 : a.N - 1;
            __syncwarp();
            if (lane < 2) u0[lane] = had ? plan[pos * 2 + lane] : a.u_prev[(size_t)b * 2 + lane];
            if (lane == 0) atomicAdd(a.n_fallback, 1ull);
        }
        if (lane == 0) { a.plan_pos[b] = pos; a.have_plan[b] = (had || ok) ? 1 : 0; }
        __syncwarp();
    }
    if (lane < 2) {
        const double u = u0[lane];
        a.u_prev[(size_t)b * 2 + lane] = u;
        a.aux[(size_t)b * 36 + lane] = u;
    }
}
}  // namespace

extern "C" int dart_lmpc_post_step(int32_t B, int32_t N, const int32_t* status, const uint8_t* fresh, const double* w_prev,
                                   double* w_next, double* u0, double* u_prev, double* aux, double* plan_U, int64_t* plan_pos,
                                   uint8_t* have_plan, uint64_t* n_fallback, void* stream) {
    if (B < 0 || N < 1 || !u0 || !u_prev || !aux) return DART_ERR_ARG;
    if (plan_U && (!status || !w_prev || !w_next || !plan_pos || !have_plan || !n_fallback)) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    PostArgs a{B, N, LmpcAxis::nw(N), status, fresh, w_prev, w_next, u0, u_prev, aux, plan_U, (long long*)plan_pos, have_plan,
               (unsigned long long*)n_fallback};
    const long threads = (long)B * 32;
    lmpc_post_step_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_lmpc_plant_step(int32_t B, double Ts, const double* true_aux, const double* u, const double* state,
                                    double* state_out, void* stream) {
    if (B < 0 || !(Ts > 0.0) || !true_aux || !u || !state || !state_out) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    KArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.cfg.Ts = Ts;
    a.x0 = state; a.ref = state; a.aux = true_aux;        // load() reads the reference row too; its value is not used here
    lmpc_plant_step_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, u, state_out);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}
