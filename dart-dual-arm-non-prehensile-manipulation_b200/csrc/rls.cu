// RMPC online adaptation on the device.
//
//   rls_update_kernel      RLS.update (np_mpc_adaptive_with_linear_regressor.py:17-27), one thread per
//                          (instance, estimator); theta (7) and P (7x7) live in registers for the update.
//                          Operation order is the reference's: denom = lam + (phi P) phi; K = (P phi)/denom;
//                          err = y - phi.theta; theta += K err; P = (P - outer(K,phi) P)/lam  (not symmetrised).
//   rmpc_prologue_kernel   one closed-loop step's host glue of RMPC/dev_dual/rob_ctrl.py:335-351 fused in one
//                          launch: a_meas = (v - v_prev)/Ts, phi(prev_state), two RLS updates, reference governor
//                          r_v += alpha*clip(target - r_v, +-dr_max), build_ref_traj, and the solver's aux row
//                          [u_prev, theta_hat].
#include <cuda_runtime.h>
#include "../../include/dart_b200.h"

namespace {

constexpr int P7 = 7;

__device__ __forceinline__ void rls_update_regs(double* th, double* Pm, const double* phi, double y, double lam) {
    double v[P7], u[P7];
    // v = phi @ P (row vector times matrix), u = P @ phi
#pragma unroll
    for (int j = 0; j < P7; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < P7; ++i) acc += phi[i] * Pm[i * P7 + j];
        v[j] = acc;
    }
    double denom = 0.0;
#pragma unroll
    for (int j = 0; j < P7; ++j) denom += v[j] * phi[j];
    denom = lam + denom;
#pragma unroll
    for (int i = 0; i < P7; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < P7; ++j) acc += Pm[i * P7 + j] * phi[j];
        u[i] = acc;
    }
    double K[P7];
#pragma unroll
    for (int i = 0; i < P7; ++i) K[i] = u[i] / denom;
    double pt = 0.0;
#pragma unroll
    for (int i = 0; i < P7; ++i) pt += phi[i] * th[i];
    const double err = y - pt;
#pragma unroll
    for (int i = 0; i < P7; ++i) th[i] = th[i] + K[i] * err;
    // P = (P - outer(K, phi) @ P) / lam, with the outer product formed first as the reference does
    double Pn[P7 * P7];
#pragma unroll
    for (int i = 0; i < P7; ++i) {
#pragma unroll
        for (int j = 0; j < P7; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < P7; ++l) acc += (K[i] * phi[l]) * Pm[l * P7 + j];
            Pn[i * P7 + j] = (Pm[i * P7 + j] - acc) / lam;
        }
    }
#pragma unroll
    for (int i = 0; i < P7 * P7; ++i) Pm[i] = Pn[i];
}

__global__ void __launch_bounds__(128) rls_update_kernel(int nest, int E, double* __restrict__ theta, double* __restrict__ Pg,
                                                        const double* __restrict__ phi, const double* __restrict__ y, double lam) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;       // estimator index = b * E + est
    if (e >= nest) return;
    const int b = e / E;
    double th[P7], Pm[P7 * P7], ph[P7];
#pragma unroll
    for (int i = 0; i < P7; ++i) { th[i] = theta[(size_t)e * P7 + i]; ph[i] = phi[(size_t)b * P7 + i]; }
#pragma unroll
    for (int i = 0; i < P7 * P7; ++i) Pm[i] = Pg[(size_t)e * P7 * P7 + i];
    rls_update_regs(th, Pm, ph, y[e], lam);
#pragma unroll
    for (int i = 0; i < P7; ++i) theta[(size_t)e * P7 + i] = th[i];
#pragma unroll
    for (int i = 0; i < P7 * P7; ++i) Pg[(size_t)e * P7 * P7 + i] = Pm[i];
}

struct PrologueArgs {
    int B, N;
    double Ts, v_eps, lam, dr_max, alpha_rg, step_fraction;
    const double *xk, *target, *u_prev;
    double *prev_state, *r_v, *theta, *P, *ref, *aux;
};

__global__ void __launch_bounds__(128) rmpc_prologue_kernel(const PrologueArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;       // (instance, estimator) pairs, estimator 0 = x, 1 = y
    if (e >= 2 * a.B) return;
    const int b = e >> 1, est = e & 1;
    const double* xs = a.xk + (size_t)b * 4;
    const double* ps = a.prev_state + (size_t)b * 4;
    // regressor of the previous state and finite-difference acceleration (rob_ctrl.py:336-339)
    double ph[P7];
    ph[0] = ps[0]; ph[1] = ps[1]; ph[2] = ps[2]; ph[3] = ps[3];
    ph[4] = tanh(ps[1] / a.v_eps); ph[5] = tanh(ps[3] / a.v_eps); ph[6] = 1.0;
    const double ymeas = (xs[1 + 2 * est] - ps[1 + 2 * est]) / a.Ts;
    // prev_state <- xk for the next step (rob_ctrl.py:365 `prev_state = xk.copy()`): each of the instance's two threads
    // writes its own axis once BOTH have read the old values (they are neighbours in one warp)
    const double nx0 = xs[2 * est], nx1 = xs[2 * est + 1];
    __syncwarp();
    a.prev_state[(size_t)b * 4 + 2 * est] = nx0;
    a.prev_state[(size_t)b * 4 + 2 * est + 1] = nx1;
    double th[P7], Pm[P7 * P7];
#pragma unroll
    for (int i = 0; i < P7; ++i) th[i] = a.theta[(size_t)e * P7 + i];
#pragma unroll
    for (int i = 0; i < P7 * P7; ++i) Pm[i] = a.P[(size_t)e * P7 * P7 + i];
    rls_update_regs(th, Pm, ph, ymeas, a.lam);
#pragma unroll
    for (int i = 0; i < P7; ++i) { a.theta[(size_t)e * P7 + i] = th[i]; a.aux[(size_t)b * 16 + 2 + est * P7 + i] = th[i]; }
#pragma unroll
    for (int i = 0; i < P7 * P7; ++i) a.P[(size_t)e * P7 * P7 + i] = Pm[i];
    a.aux[(size_t)b * 16 + est] = a.u_prev[(size_t)b * 2 + est];
    // reference governor on this axis' position (rob_ctrl.py:346-348) and staged reference (np_mpc...:201-210)
    const int pi = 2 * est;
    const double tg = a.target[(size_t)b * 4 + pi];
    double rv = a.r_v[(size_t)b * 4 + pi];
    const double err = tg - rv;
    rv = rv + a.alpha_rg * fmin(fmax(err, -a.dr_max), a.dr_max);
    a.r_v[(size_t)b * 4 + pi] = rv;
    double* ref = a.ref + (size_t)b * (a.N + 1) * 4;
    double decay = 1.0;
    for (int i = 0; i <= a.N; ++i) {
        // w = 1 - (1 - step_fraction)^(i+1); numpy evaluates the power directly
        const double w = 1.0 - pow(1.0 - a.step_fraction, (double)(i + 1));
        ref[i * 4 + pi] = rv + w * (tg - rv);
        ref[i * 4 + pi + 1] = 0.0;
    }
    (void)decay;
}

}  // namespace

extern "C" int dart_rls_update(int32_t B, int32_t E, double* theta, double* P, const double* phi, const double* y,
                               double lam, void* stream) {
    if (B < 0 || E < 1 || !theta || !P || !phi || !y || !(lam > 0.0)) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    const int nest = B * E;
    rls_update_kernel<<<(nest + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nest, E, theta, P, phi, y, lam);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int dart_rmpc_prologue(int32_t B, int32_t N, double Ts, double v_eps, double lam, double dr_max, double alpha_rg,
                                  double step_fraction, const double* xk, double* prev_state, const double* target,
                                  const double* u_prev, double* r_v, double* theta, double* P, double* ref, double* aux,
                                  void* stream) {
    if (B < 0 || N < 1 || !xk || !prev_state || !target || !u_prev || !r_v || !theta || !P || !ref || !aux) return DART_ERR_ARG;
    if (!(Ts > 0.0) || !(v_eps > 0.0) || !(lam > 0.0)) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    PrologueArgs a{B, N, Ts, v_eps, lam, dr_max, alpha_rg, step_fraction, xk, target, u_prev, prev_state, r_v, theta, P, ref, aux};
    rmpc_prologue_kernel<<<(2 * B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}
