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
}  // namespace

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
