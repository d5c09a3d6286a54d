// Surrogate plant step + episode metrics as a device function (shared by the per-step kernel in plant.cu and the
// persistent episode kernel in nmpc_kernel.cuh).  See plant.cu for the model and the metric definitions.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace dart {

struct PlantArgs {
    int B;
    double Ts, g, tol;
    const double *mu, *coulomb, *u, *target;
    double *state, *conv_time, *effort, *err;
    int32_t* nsteps;     // per-instance step counter on the device: keeps the launch replayable from a CUDA graph
};

__device__ __forceinline__ void plant_f(const double* x, double sx, double sy, double vn, double g, double mu, double c,
                                        double Ts, double* f) {
    f[0] = x[1];
    f[1] = g * sx - mu * x[1] - c * fabs(g) * tanh(x[1] / 0.01);
    f[2] = x[3];
    f[3] = g * sy - mu * x[3] - c * fabs(g) * tanh(x[3] / 0.01);
    f[4] = vn;
    f[5] = (vn - x[5]) / Ts;
}

__device__ __forceinline__ void plant_step_one(const PlantArgs& a, int b) {
    double x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = a.state[(size_t)b * 6 + i];
    const double ux = a.u[(size_t)b * 2], uy = a.u[(size_t)b * 2 + 1];
    // metrics on the state that was logged for this step
    const double ex = x[0] - a.target[(size_t)b * 6 + 0], ey = x[2] - a.target[(size_t)b * 6 + 2];
    const double e = sqrt(ex * ex + ey * ey);
    a.err[b] = e;
    const int32_t step = a.nsteps[b];
    a.nsteps[b] = step + 1;
    if (a.conv_time[b] < 0.0 && e < a.tol) a.conv_time[b] = step * a.Ts;
    a.effort[b] += sqrt(ux * ux + uy * uy) * a.Ts;
    // RK4, input held
    const double mu = a.mu[b], c = a.coulomb ? a.coulomb[b] : 0.0, Ts = a.Ts, g = a.g;
    const double sx = sin(ux), sy = sin(uy), vn = -g * (ux * ux + uy * uy);
    double k1[6], k2[6], k3[6], k4[6], t[6];
    plant_f(x, sx, sy, vn, g, mu, c, Ts, k1);
#pragma unroll
    for (int i = 0; i < 6; ++i) t[i] = x[i] + Ts / 2 * k1[i];
    plant_f(t, sx, sy, vn, g, mu, c, Ts, k2);
#pragma unroll
    for (int i = 0; i < 6; ++i) t[i] = x[i] + Ts / 2 * k2[i];
    plant_f(t, sx, sy, vn, g, mu, c, Ts, k3);
#pragma unroll
    for (int i = 0; i < 6; ++i) t[i] = x[i] + Ts * k3[i];
    plant_f(t, sx, sy, vn, g, mu, c, Ts, k4);
#pragma unroll
    for (int i = 0; i < 6; ++i) a.state[(size_t)b * 6 + i] = x[i] + Ts / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
}

}  // namespace dart
