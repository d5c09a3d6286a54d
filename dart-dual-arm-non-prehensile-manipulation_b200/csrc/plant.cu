// Surrogate closed loop on the device: plant step + the reference logger's episode metrics, so whole
// 18-object x many-target episodes run without host round trips (solve kernel -> this kernel, per simulated step).
//
// Plant = the PMPC model itself (mpc_3d.py:87-104, literal RK4 with the tilt held over Ts) plus an optional Coulomb
// term  -c*|g|*tanh(v/0.01)  that the controller's model does not know.  Metrics follow PMPC/src/logger.py:155-176:
// steady-state error = |pos - target| of the last logged state, convergence time = first logged time with error
// < tol (else the final time), control effort = sum |u| dt.
#include <cuda_runtime.h>
#include "../../include/dart_b200.h"
#include "plant.cuh"

namespace {
using dart::PlantArgs;

// status / iters / counters (all or none): the solve statistics of the step that produced `u` are accumulated here as
// the persistent episode kernel does (counters[0] += iterations, counters[1] += solves that did not end converged), so the
// stepwise closed loop needs no reduction launches of its own
__global__ void __launch_bounds__(128) pmpc_plant_step_kernel(const PlantArgs a, const int32_t* __restrict__ status,
                                                              const int32_t* __restrict__ iters, unsigned long long* counters) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = b < a.B;
    if (counters != nullptr) {
        unsigned long long it = in ? (unsigned long long)iters[b] : 0ull;
        unsigned nc = (in && status[b] != 0 /* ST_CONVERGED */) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { it += __shfl_xor_sync(0xffffffffu, it, o); nc += __shfl_xor_sync(0xffffffffu, nc, o); }
        if ((threadIdx.x & 31) == 0) {
            if (it) atomicAdd(&counters[0], it);
            if (nc) atomicAdd(&counters[1], (unsigned long long)nc);
        }
    }
    if (!in) return;
    dart::plant_step_one(a, b);
}

// Surrogate plant of BASELINE config 3: v' = gz sin(u) - mu |g| tanh(v / 0.01) - c v per axis, four explicit
// Euler sub-steps of Ts / 4 (positions first, with the old velocity) -- workloads.rmpc_plant_step on the host.
__global__ void __launch_bounds__(128) rmpc_plant_step_kernel(int B, double Ts, double gz, const double* __restrict__ mu,
                                                              const double* __restrict__ cdamp, const double* __restrict__ u,
                                                              const double* __restrict__ x_in, double* __restrict__ x_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double px = x_in[b * 4 + 0], vx = x_in[b * 4 + 1], py = x_in[b * 4 + 2], vy = x_in[b * 4 + 3];
    const double sx = gz * sin(u[b * 2 + 0]), sy = gz * sin(u[b * 2 + 1]);
    const double m = mu[b] * 9.81, c = cdamp[b], h = Ts / 4;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const double ax = sx - m * tanh(vx / 0.01) - c * vx;
        const double ay = sy - m * tanh(vy / 0.01) - c * vy;
        px += h * vx; py += h * vy;
        vx += h * ax; vy += h * ay;
    }
    x_out[b * 4 + 0] = px; x_out[b * 4 + 1] = vx; x_out[b * 4 + 2] = py; x_out[b * 4 + 3] = vy;
}

}  // namespace

extern "C" int dart_rmpc_plant_step(int32_t B, double Ts, double gz, const double* mu, const double* c, const double* u,
                                    const double* state, double* state_out, void* stream) {
    if (B < 0 || !mu || !c || !u || !state || !state_out || !(Ts > 0.0)) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    rmpc_plant_step_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, Ts, gz, mu, c, u, state, state_out);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_pmpc_plant_step(int32_t B, double Ts, double g, const double* mu, const double* coulomb, const double* u,
                                    const double* target, double* state, int32_t* nsteps, double tol, double* conv_time,
                                    double* effort, double* err, const int32_t* status, const int32_t* iters,
                                    uint64_t* counters, void* stream) {
    if (B < 0 || !mu || !u || !target || !state || !nsteps || !conv_time || !effort || !err || !(Ts > 0.0)) return DART_ERR_ARG;
    if (counters && (!status || !iters)) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    PlantArgs a{B, Ts, g, tol, mu, coulomb, u, target, state, conv_time, effort, err, nsteps};
    pmpc_plant_step_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, status, iters, (unsigned long long*)counters);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}
