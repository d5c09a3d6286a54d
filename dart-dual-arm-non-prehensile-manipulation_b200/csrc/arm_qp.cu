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
#include <cmath>
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

// =====================================================================================================================
// QP build on the device: arm.py:337-405 for B arms, from the dictionary ARMCONTROL.compute_dynamics returns.  The
// reference does this in numpy per cycle (pinv of the mass matrix, inverse or pinv of the task-space inertia, its matrix
// square root by eigh); here both symmetric eigen-decompositions are cyclic Jacobi iterations run by the 8 lanes of the
// arm's tile (lane k owns row/column k of the rotation), and every matrix function follows from them:
//   pinv(M, rcond 1e-6)               = V7 diag(1/w | 0) V7'
//   Mx = inv(Mx_inv) or pinv(., 1e-3) = V6 diag(1/w | 0) V6'      (arm.py:351-357 picks by |det| > 1e-8)
//   safe_matrix_sqrt(Mx)              = V6 diag(sqrt|1/w|) V6'    (arm.py:363-366)
namespace dart {
namespace {

struct ArmParams {
    double Wimp[36], Wpos[49], Wsm[49], K[36], sK[36], Knull[49], lim_lo[21], lim_hi[21], dt;      // sK = sqrt(K) elementwise
};

struct BuildSmem {
    double A7[49], V7[49], A6[36], V6[36], Mx[36], sMx[36], w7[7], w6[6], sw6[6];
    double qd[7], q[7], jq[6], jdq[6], v1[7], v2[6], F[6], e0[6], beta[7], we0[6];
};
constexpr int kBuildDoubles = sizeof(BuildSmem) / sizeof(double) | 1;

// Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix, ONE THREAD per matrix: the upper triangle lives in
// registers (all indices are compile-time after unrolling), the eigenvector matrix V in the thread's shared-memory
// block.  A warp thus advances 32 arms per instruction and needs no synchronisation -- the 8-lanes-per-arm version of
// this routine spent 60 % of the build kernel in rotation arithmetic replicated across lanes and in tile syncs.
// On return the diagonal of A (shared memory) holds the eigenvalues and the columns of V the eigenvectors.
template <int n>
__device__ __forceinline__ void jacobi_serial(double* A, double* V) {
    double a[n * n];                                         // a[i * n + j], i <= j used
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = i; j < n; ++j) a[i * n + j] = (i == j) ? A[i * n + i] : 0.5 * (A[i * n + j] + A[j * n + i]);
#pragma unroll
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0, dia = 0.0;
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = i; j < n; ++j) { if (i == j) dia += a[i * n + j] * a[i * n + j]; else off += a[i * n + j] * a[i * n + j]; }
        if (2.0 * off <= 1e-32 * dia) break;
#pragma unroll
        for (int p = 0; p < n - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p * n + q];
                if (fabs(apq) > 1e-300) {
                    const double d = a[q * n + q] - a[p * n + p], b2 = 2.0 * apq;
                    const double t = (d >= 0.0 ? b2 : -b2) / (fabs(d) + sqrt(d * d + b2 * b2));
                    const double c = rsqrt(t * t + 1.0), s = t * c;
                    a[p * n + p] -= t * apq;
                    a[q * n + q] += t * apq;
                    a[p * n + q] = 0.0;
#pragma unroll
                    for (int k = 0; k < n; ++k) {
                        if (k == p || k == q) continue;
                        const int ip = (k < p) ? k * n + p : p * n + k, iq = (k < q) ? k * n + q : q * n + k;
                        const double akp = a[ip], akq = a[iq];
                        a[ip] = c * akp - s * akq;
                        a[iq] = s * akp + c * akq;
                    }
#pragma unroll
                    for (int k = 0; k < n; ++k) {
                        const double vkp = V[k * n + p], vkq = V[k * n + q];
                        V[k * n + p] = c * vkp - s * vkq;
                        V[k * n + q] = s * vkp + c * vkq;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < n; ++i) A[i * n + i] = a[i * n + i];
}

__global__ void __launch_bounds__(256) arm_qp_build_kernel(int B, ArmParams P, const double* __restrict__ q_, const double* __restrict__ qd_,
                                                            const double* __restrict__ qddp_, const double* __restrict__ mocap_,
                                                            const double* __restrict__ ee_, const double* __restrict__ rotvec_,
                                                            const double* __restrict__ jac_, const double* __restrict__ jacDot_,
                                                            const double* __restrict__ M_, const double* __restrict__ h_,
                                                            const double* __restrict__ Mxinv_, double* __restrict__ Hout,
                                                            double* __restrict__ gout, double* __restrict__ c0out,
                                                            double* __restrict__ Cout, double* __restrict__ loout,
                                                            double* __restrict__ hiout) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & (G - 1);
    const unsigned mask = 0xffu << ((threadIdx.x & 31) - lane);
    const int qib = threadIdx.x / G;
    const int apb = blockDim.x / G;                          // arms per block (32)
    const long b0 = (long)blockIdx.x * apb;
    const long b = b0 + qib;
    const bool active = b < B;
    BuildSmem& S = *reinterpret_cast<BuildSmem*>(smem + (size_t)qib * kBuildDoubles);
    const double* jac = jac_ + b * 42;
    const double* jacDot = jacDot_ + b * 42;
    const double* Mg = M_ + b * 49;
    const double* Mxi = Mxinv_ + b * 36;
    if (active) {
        for (int i = lane; i < 49; i += G) S.A7[i] = Mg[i];
        for (int i = lane; i < 36; i += G) S.A6[i] = Mxi[i];
        if (lane < 7) { S.qd[lane] = qd_[b * 7 + lane]; S.q[lane] = q_[b * 7 + lane]; }
    }
    __syncthreads();
    // ---- eigen-decompositions, one thread per matrix: warp 0 takes the mass matrices, warp 1 the task-space inertias
    if (threadIdx.x < 2 * apb) {
        const int arm = threadIdx.x % apb;
        if (b0 + arm < B) {
            BuildSmem& T = *reinterpret_cast<BuildSmem*>(smem + (size_t)arm * kBuildDoubles);
            if (threadIdx.x < apb) jacobi_serial<7>(T.A7, T.V7);
            else jacobi_serial<6>(T.A6, T.V6);
        }
    }
    __syncthreads();
    if (!active) return;
    // ---- spectra -> reciprocal spectra (pinv thresholds on singular values = |eigenvalues|)
    {
        double m7 = 0.0, m6 = 0.0, det = 1.0;
        for (int k = 0; k < 7; ++k) m7 = fmax(m7, fabs(S.A7[k * 7 + k]));
        for (int k = 0; k < 6; ++k) { m6 = fmax(m6, fabs(S.A6[k * 6 + k])); det *= S.A6[k * 6 + k]; }
        const bool regular = fabs(det) > 1e-8;
        __syncwarp(mask);
        if (lane < 7) { const double w = S.A7[lane * 7 + lane]; S.w7[lane] = (fabs(w) > 1e-6 * m7) ? 1.0 / w : 0.0; }
        if (lane < 6) {
            const double w = S.A6[lane * 6 + lane];
            const double iw = (regular || fabs(w) > 1e-3 * m6) ? 1.0 / w : 0.0;
            S.w6[lane] = iw; S.sw6[lane] = sqrt(fabs(iw));
        }
    }
    __syncwarp(mask);
    // ---- Mx, sqrt(Mx) (column `lane`), v1 = pinv(M) h, J qd, Jdot qd
    if (lane < 6) {
        for (int i = 0; i < 6; ++i) {
            double a = 0.0, r = 0.0;
            for (int k = 0; k < 6; ++k) {
                const double vv = S.V6[i * 6 + k] * S.V6[lane * 6 + k];
                a += vv * S.w6[k]; r += vv * S.sw6[k];
            }
            S.Mx[i * 6 + lane] = a; S.sMx[i * 6 + lane] = r;
        }
        double a = 0.0, d = 0.0;
        for (int v = 0; v < 7; ++v) { a += jac[lane * 7 + v] * S.qd[v]; d += jacDot[lane * 7 + v] * S.qd[v]; }
        S.jq[lane] = a; S.jdq[lane] = d;
    }
    if (lane < 7) {
        double acc = 0.0;                                    // (V diag(w) V' h)_lane
        for (int k = 0; k < 7; ++k) {
            double vh = 0.0;
            for (int i = 0; i < 7; ++i) vh += S.V7[i * 7 + k] * h_[b * 7 + i];
            acc += S.V7[lane * 7 + k] * S.w7[k] * vh;
        }
        S.v1[lane] = acc;
        double be = 2.0 * sqrt(P.Knull[lane * 7 + lane]) * (-S.qd[lane]);
        for (int v = 0; v < 7; ++v) be += P.Knull[lane * 7 + v] * (-S.q[v]);
        S.beta[lane] = be;
    }
    __syncwarp(mask);
    if (lane < 6) {
        double a = S.jdq[lane];
        for (int v = 0; v < 7; ++v) a += jac[lane * 7 + v] * S.v1[v];
        S.v2[lane] = a;
    }
    __syncwarp(mask);
    // ---- F = -D (J qd) + K twist + Mx v2,  D = sqrt(Mx) sqrt(K) + sqrt(K) sqrt(Mx)  (sqrt(K) elementwise, arm.py:368)
    if (lane < 6) {
        double f = 0.0;
        for (int c = 0; c < 6; ++c) {
            double d = 0.0;
            for (int k = 0; k < 6; ++k) d += S.sMx[lane * 6 + k] * P.sK[k * 6 + c] + P.sK[lane * 6 + k] * S.sMx[k * 6 + c];
            const double tw = (c < 3) ? (mocap_[b * 3 + c] - ee_[b * 3 + c]) : rotvec_[b * 3 + (c - 3)];
            f += -d * S.jq[c] + P.K[lane * 6 + c] * tw + S.Mx[lane * 6 + c] * S.v2[c];
        }
        S.F[lane] = f;
    }
    __syncwarp(mask);
    if (lane < 6) {
        double e = S.jdq[lane];
        for (int c = 0; c < 6; ++c) e -= Mxi[lane * 6 + c] * S.F[c];
        S.e0[lane] = e;
    }
    __syncwarp(mask);
    if (lane < 6) {
        double a = 0.0;
        for (int c = 0; c < 6; ++c) a += P.Wimp[lane * 6 + c] * S.e0[c];
        S.we0[lane] = a;
    }
    __syncwarp(mask);
    // ---- H column `lane`, g, bounds, C; c0 by reduction
    double c0p = 0.0;
    if (lane < 7) {
        double wj[6];                                        // (Wimp J)[:, lane]
        for (int r = 0; r < 6; ++r) { double a = 0.0; for (int c = 0; c < 6; ++c) a += P.Wimp[r * 6 + c] * jac[c * 7 + lane]; wj[r] = a; }
        for (int i = 0; i < 7; ++i) {
            double a = 0.0, at = 0.0;                        // J'WJ [i][lane] and its transpose entry (Wimp may be non-symmetric)
            for (int r = 0; r < 6; ++r) a += jac[r * 7 + i] * wj[r];
            for (int r = 0; r < 6; ++r) { double w2 = 0.0; for (int c = 0; c < 6; ++c) w2 += P.Wimp[r * 6 + c] * jac[c * 7 + i]; at += jac[r * 7 + lane] * w2; }
            const double hij = 2.0 * (a + P.Wpos[i * 7 + lane] + P.Wsm[i * 7 + lane]);
            const double hji = 2.0 * (at + P.Wpos[lane * 7 + i] + P.Wsm[lane * 7 + i]);
            Hout[b * 49 + i * 7 + lane] = 0.5 * (hij + hji);
        }
        double gg = 0.0, wb = 0.0, ws = 0.0;
        for (int r = 0; r < 6; ++r) gg += jac[r * 7 + lane] * S.we0[r];
        for (int v = 0; v < 7; ++v) { wb += P.Wpos[lane * 7 + v] * S.beta[v]; ws += P.Wsm[lane * 7 + v] * qddp_[b * 7 + v]; }
        gout[b * 7 + lane] = 2.0 * (gg - wb - ws);
        c0p = S.beta[lane] * wb + qddp_[b * 7 + lane] * ws;
        const double qd = S.qd[lane], q = S.q[lane], hh = h_[b * 7 + lane];
        const double off[3] = {qd * P.dt + q, qd, hh};
        for (int t = 0; t < 3; ++t) {
            loout[b * 21 + t * 7 + lane] = P.lim_lo[t * 7 + lane] - off[t];
            hiout[b * 21 + t * 7 + lane] = P.lim_hi[t * 7 + lane] - off[t];
        }
        for (int v = 0; v < 7; ++v) {
            Cout[b * 147 + lane * 7 + v] = (v == lane) ? 0.5 * P.dt * P.dt : 0.0;
            Cout[b * 147 + (7 + lane) * 7 + v] = (v == lane) ? P.dt : 0.0;
            Cout[b * 147 + (14 + lane) * 7 + v] = Mg[lane * 7 + v];
        }
    }
    if (lane < 6) c0p += S.e0[lane] * S.we0[lane];
    c0p = tsum(c0p, mask);
    if (lane == 0) c0out[b] = c0p;
}

}  // namespace
}  // namespace dart

extern "C" int dart_arm_qp_build(int32_t B, const double* Wimp, const double* Wpos, const double* Wsmooth, const double* K,
                                 const double* K_null, const double* limits_lo, const double* limits_hi, double dt,
                                 const double* q, const double* qd, const double* qdd_prev, const double* mocap_pos,
                                 const double* ee_pos, const double* rotvec, const double* jac, const double* jacDot,
                                 const double* M, const double* h, const double* Mx_inv, double* H, double* g, double* c0,
                                 double* C, double* lo, double* hi, void* stream) {
    if (B < 0 || !Wimp || !Wpos || !Wsmooth || !K || !K_null || !limits_lo || !limits_hi || !(dt > 0) || !q || !qd || !qdd_prev ||
        !mocap_pos || !ee_pos || !rotvec || !jac || !jacDot || !M || !h || !Mx_inv || !H || !g || !c0 || !C || !lo || !hi)
        return DART_ERR_ARG;
    if (B == 0) return DART_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DART_ERR_NO_DEVICE;
    dart::ArmParams P;
    for (int i = 0; i < 36; ++i) { P.Wimp[i] = Wimp[i]; P.K[i] = K[i]; P.sK[i] = sqrt(K[i]); }
    for (int i = 0; i < 49; ++i) { P.Wpos[i] = Wpos[i]; P.Wsm[i] = Wsmooth[i] / (dt * dt); P.Knull[i] = K_null[i]; }
    for (int i = 0; i < 21; ++i) { P.lim_lo[i] = limits_lo[i]; P.lim_hi[i] = limits_hi[i]; }
    P.dt = dt;
    const int threads = 256, qpb = threads / dart::G;
    const size_t smem = (size_t)qpb * dart::kBuildDoubles * sizeof(double);
    static bool attr_set[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return DART_ERR_CUDA;
    if (!attr_set[dev]) {
        if (cudaFuncSetAttribute(dart::arm_qp_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return DART_ERR_CUDA;
        attr_set[dev] = true;
    }
    dart::arm_qp_build_kernel<<<(B + qpb - 1) / qpb, threads, smem, (cudaStream_t)stream>>>(
        B, P, q, qd, qdd_prev, mocap_pos, ee_pos, rotvec, jac, jacDot, M, h, Mx_inv, H, g, c0, C, lo, hi);
    ++g_arm_qp_launches;
    return cudaGetLastError() == cudaSuccess ? DART_OK : DART_ERR_CUDA;
}
