// Batched low-level arm QP: the 7-variable, 21-row problem that PMPC/src/controller/arm.py:337-457 rebuilds
// and hands to CasADi/IPOPT every 2 ms per arm,
//     min 0.5 x'Hx + g'x   s.t.  lo <= Cx <= hi,      x = joint accelerations,
// solved by a primal-dual interior-point method in slack form (Cx - s = 0, lo <= s <= hi; monotone barrier, fraction to
// the boundary; the problem is strictly convex, so no line search is needed).
//
// Mapping: a tile of 8 lanes per QP, 4 QPs per warp.  Lane j < 7 owns variable j and rows {j, 7+j, 14+j} -- in the
// reference those are its position bound, its velocity bound and its torque row -- whose slacks and multipliers stay in
// that lane's registers for the whole solve.  H, C and the vectors the lanes exchange live in a per-QP shared-memory
// block.  Per iteration: row residuals (lane-local) -> dual residual and KKT column j = H[:,j] + C' Sigma C[:,j]
// (lane-local, 147 FMAs) -> every lane factors the 7x7 system from shared memory in registers (no exchange) ->
// lane-local step lengths -> tile reductions by shuffle.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/dart_b200.h"

namespace dart {
namespace {

constexpr int NV = 7, NR = 21, G = 8, RPL = 3;      // variables, rows, lanes per QP, rows per lane

struct QpSmem {
    double H[NV * NV], C[NR * NV], g[NV], x[NV], nu[NR], sig[NR], nh[NR], K[NV * NV], rhs[NV];
};
constexpr int kQpDoubles = sizeof(QpSmem) / sizeof(double) | 1;      // odd stride: the 4 QPs of a warp on distinct banks

__device__ __forceinline__ double tmax(double v, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o, G));
    return v;
}
__device__ __forceinline__ double tmin(double v, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o, G));
    return v;
}
__device__ __forceinline__ double tsum(double v, unsigned mask) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
    return v;
}

__global__ void __launch_bounds__(128) arm_qp_kernel(int B, const double* __restrict__ Hg, const double* __restrict__ gg,
                                                      const double* __restrict__ Cg, const double* __restrict__ log_,
                                                      const double* __restrict__ hig, const double* __restrict__ x0g,
                                                      double* __restrict__ xg, double* __restrict__ objg,
                                                      int32_t* __restrict__ stg, int32_t* __restrict__ itg, double tol,
                                                      int max_iter) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & (G - 1);
    const unsigned mask = 0xffu << ((threadIdx.x & 31) - lane);
    const int qib = threadIdx.x / G;
    const long qp = (long)blockIdx.x * (blockDim.x / G) + qib;
    if (qp >= B) return;                                   // whole tiles leave together (B is checked per tile)
    QpSmem& S = *reinterpret_cast<QpSmem*>(smem + (size_t)qib * kQpDoubles);
    const bool own = lane < NV;
    const int j = own ? lane : NV - 1;                     // lane 7 shadows lane 6 (no stores)

    // ---- load the problem: coalesced over the tile
    for (int i = lane; i < NV * NV; i += G) S.H[i] = Hg[qp * NV * NV + i];
    for (int i = lane; i < NR * NV; i += G) S.C[i] = Cg[qp * NR * NV + i];
    if (own) { S.g[lane] = gg[qp * NV + lane]; S.x[lane] = x0g ? x0g[qp * NV + lane] : 0.0; }
    __syncwarp(mask);

    const double mu0 = 0.1, kappa_mu = 0.2, kappa_eps = 10.0, tau_min = 0.99, bound_push = 1e-2, smax = 100.0;
    const double mu_min = tol / 10.0;
    double lo[RPL], hi[RPL], s[RPL], zl[RPL], zu[RPL];
    bool empty = false;
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int r = q * NV + j;
        lo[q] = log_[qp * NR + r]; hi[q] = hig[qp * NR + r];
        double t = 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) t += S.C[r * NV + v] * S.x[v];
        const double push = fmin(bound_push * fmax(1.0, fmax(fabs(lo[q]), fabs(hi[q]))), bound_push * (hi[q] - lo[q]));
        s[q] = fmin(fmax(t, lo[q] + push), hi[q] - push);
        zl[q] = mu0 / (s[q] - lo[q]);
        zu[q] = mu0 / (hi[q] - s[q]);
        empty |= !(hi[q] - lo[q] > 0.0);
    }
    empty = __any_sync(mask, empty);
    double mu = mu0;
    int it = 0;
    int32_t st = empty ? 2 : 1;                            // DART_STATUS_INFEASIBLE : DART_STATUS_MAXITER
    bool done = empty;

    while (!done) {
        // ---- A: row residuals, complementarity
        double rc[RPL], isl[RPL], isu[RPL];
        double pinf = 0.0, zmn = 1e300, zmx = 0.0, zs = 0.0;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int r = q * NV + j;
            double t = 0.0;
#pragma unroll
            for (int v = 0; v < NV; ++v) t += S.C[r * NV + v] * S.x[v];
            rc[q] = t - s[q];
            const double sl = s[q] - lo[q], su = hi[q] - s[q];
            const double ip = 1.0 / (sl * su);
            isl[q] = su * ip; isu[q] = sl * ip;
            const double a = zl[q] * sl, b = zu[q] * su;
            pinf = fmax(pinf, fabs(rc[q]));
            zmn = fmin(zmn, fmin(a, b)); zmx = fmax(zmx, fmax(a, b));
            zs += fabs(zl[q]) + fabs(zu[q]);
            if (own) S.nu[r] = zu[q] - zl[q];
        }
        if (!own) { zs = 0.0; }
        __syncwarp(mask);
        // ---- B: dual residual
        double grad = S.g[j];
#pragma unroll
        for (int v = 0; v < NV; ++v) grad += S.H[j * NV + v] * S.x[v];
        double dres = grad;
#pragma unroll
        for (int r = 0; r < NR; ++r) dres += S.C[r * NV + j] * S.nu[r];
        const double dinf = tmax(fabs(dres), mask);
        pinf = tmax(pinf, mask); zmx = tmax(zmx, mask); zmn = tmin(zmn, mask); zs = tsum(zs, mask);
        const double mz = zs / (2.0 * NR);
        const double isd = (mz > smax) ? smax / mz : 1.0;
        const double base = fmax(dinf * isd, pinf);
        const double E0 = fmax(base, zmx * isd);
        if (E0 <= tol) { st = 0; break; }
        if (!(E0 == E0) || E0 > 1e300) { st = 3; break; }
        if (it >= max_iter) break;
        ++it;
        for (int q = 0; q < 8; ++q) {
            const double Emu = fmax(base, fmax(fabs(zmx - mu), fabs(zmn - mu)) * isd);
            if (Emu <= kappa_eps * mu && mu > mu_min) mu = fmax(mu_min, fmin(kappa_mu * mu, mu * sqrt(mu)));
            else break;
        }
        // ---- C: barrier curvature and condensed multipliers of the own rows
        double sig[RPL];
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int r = q * NV + j;
            sig[q] = zl[q] * isl[q] + zu[q] * isu[q];
            if (own) { S.sig[r] = sig[q]; S.nh[r] = mu * (isu[q] - isl[q]) + sig[q] * rc[q]; }
        }
        __syncwarp(mask);
        // ---- D: KKT column j and right-hand side
        {
            double col[NV], rh = grad;
#pragma unroll
            for (int i = 0; i < NV; ++i) col[i] = S.H[i * NV + j];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const double cj = S.C[r * NV + j];
                const double w = S.sig[r] * cj;
                rh += cj * S.nh[r];
#pragma unroll
                for (int i = 0; i < NV; ++i) col[i] += w * S.C[r * NV + i];
            }
            if (own) {
#pragma unroll
                for (int i = 0; i < NV; ++i) S.K[i * NV + j] = col[i];
                S.rhs[j] = -rh;
            }
        }
        __syncwarp(mask);
        // ---- E: every lane factors K = L L' (lower triangle) and solves, in registers
        double L[NV * (NV + 1) / 2], dx[NV];
        bool bad = false;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
#pragma unroll
            for (int r = c; r < NV; ++r) {
                double v = S.K[r * NV + c];
#pragma unroll
                for (int k = 0; k < c; ++k) v -= L[r * (r + 1) / 2 + k] * L[c * (c + 1) / 2 + k];
                if (r == c) { bad |= !(v > 0.0); L[c * (c + 1) / 2 + c] = rsqrt(v); }      // stores 1 / L_cc
                else L[r * (r + 1) / 2 + c] = v * L[c * (c + 1) / 2 + c];
            }
        }
        if (bad) { st = 3; break; }
#pragma unroll
        for (int r = 0; r < NV; ++r) {
            double v = S.rhs[r];
#pragma unroll
            for (int k = 0; k < r; ++k) v -= L[r * (r + 1) / 2 + k] * dx[k];
            dx[r] = v * L[r * (r + 1) / 2 + r];
        }
#pragma unroll
        for (int r = NV - 1; r >= 0; --r) {
            double v = dx[r];
#pragma unroll
            for (int k = r + 1; k < NV; ++k) v -= L[k * (k + 1) / 2 + r] * dx[k];
            dx[r] = v * L[r * (r + 1) / 2 + r];
        }
        // ---- F: slack / multiplier steps of the own rows, fraction to the boundary
        double ds[RPL], dzl[RPL], dzu[RPL], rp = 0.0, rd = 0.0;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int r = q * NV + j;
            double t = rc[q];
#pragma unroll
            for (int v = 0; v < NV; ++v) t += S.C[r * NV + v] * dx[v];
            ds[q] = t;
            dzl[q] = mu * isl[q] - zl[q] - zl[q] * isl[q] * t;
            dzu[q] = mu * isu[q] - zu[q] + zu[q] * isu[q] * t;
            rp = fmax(rp, fmax(-t * isl[q], t * isu[q]));
            const double iz = 1.0 / (zl[q] * zu[q]);
            rd = fmax(rd, fmax(-dzl[q] * zu[q] * iz, -dzu[q] * zl[q] * iz));
        }
        rp = tmax(rp, mask); rd = tmax(rd, mask);
        const double tau = fmax(tau_min, 1.0 - mu);
        const double ap = (rp > tau) ? tau / rp : 1.0, ad = (rd > tau) ? tau / rd : 1.0;
        // ---- G: move
#pragma unroll
        for (int q = 0; q < RPL; ++q) { s[q] += ap * ds[q]; zl[q] += ad * dzl[q]; zu[q] += ad * dzu[q]; }
        if (own) S.x[lane] += ap * dx[lane];
        __syncwarp(mask);
    }
    // ---- result
    double hx = 0.0;
#pragma unroll
    for (int v = 0; v < NV; ++v) hx += S.H[j * NV + v] * S.x[v];
    double part = own ? S.x[j] * (0.5 * hx + S.g[j]) : 0.0;
    part = tsum(part, mask);
    if (own) xg[qp * NV + lane] = S.x[lane];
    if (lane == 0) {
        objg[qp] = part;
        if (stg) stg[qp] = st;
        if (itg) itg[qp] = it;
    }
}

}  // namespace
}  // namespace dart

static int64_t g_arm_qp_launches = 0;

extern "C" int dart_arm_qp_solve(int32_t B, const double* H, const double* g, const double* C, const double* lo,
                                 const double* hi, const double* x0, double* x, double* obj, int32_t* status,
                                 int32_t* iters, double tol, int32_t max_iter, void* stream) {
    if (B < 0 || !H || !g || !C || !lo || !hi || !x || !obj) return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DART_ERR_NO_DEVICE;
    const int threads = 128, qpb = threads / dart::G;
    const size_t smem = (size_t)qpb * dart::kQpDoubles * sizeof(double);
    static bool attr_set[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return DART_ERR_CUDA;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(dart::arm_qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return DART_ERR_CUDA;
        attr_set[dev] = true;
    }
    const int grid = (B + qpb - 1) / qpb;
    dart::arm_qp_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(B, H, g, C, lo, hi, x0, x, obj, status, iters,
                                                                       tol > 0 ? tol : 1e-8, max_iter > 0 ? max_iter : 100);
    ++g_arm_qp_launches;
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}

extern "C" int64_t dart_arm_qp_launch_count(void) { return g_arm_qp_launches; }
