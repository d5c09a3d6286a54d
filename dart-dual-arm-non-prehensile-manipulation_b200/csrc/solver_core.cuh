// Batched tray-tilt NMPC: primal-dual interior-point method with a Riccati KKT factorisation.
//
// One *problem* (a whole NLP, or one decoupled axis of it) is solved by a tile of G lanes of a warp.
// Horizon data lives in a per-problem shared-memory workspace; lanes split the horizon for everything
// that is parallel over stages (RK4 + analytic Jacobians, costs, barrier terms, KKT residuals) and split
// the columns of the stage matrices [A B d] for the serial Riccati sweep.  Tile-wide reductions are
// warp shuffles.  The same code compiles for the host with a 1-lane tile (tests/hostemu), which is how the
// solver logic is unit-tested on machines without a GPU; the product never runs it on the CPU.
//
// Replaces ca.nlpsol('ipopt') + MUMPS at the reference call sites
//   PMPC/src/controller/mpc_3d.py:82,124-132
//   RMPC/dev_dual/controller/np_mpc_adaptive_with_linear_regressor.py:157-162,214-215
//   LMPC/src/controller/rlmpc2.py:480-491,508-515
// Algorithm (IPOPT-flavoured, Waechter & Biegler 2006): slack form for every inequality row,
// monotone barrier schedule, fraction-to-boundary rule, filter-type step acceptance.  Multiple shooting
// as in the reference (states are decision variables; x_0 is eliminated).  Hessian = constant cost Hessian
// + barrier terms + the Lagrangian curvature of the tilt input, -tan(u_i) (B^T lambda)_i, which is exact
// for models that are affine in g*sin(u_i) (PMPC) and O(Ts^2)-accurate for RMPC/LMPC.  First derivatives
// are exact (forward sensitivities through RK4), so the fixed point is the NLP's KKT point.
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef DART_TRACE
#include <stdio.h>
#endif

#if defined(__CUDACC__)
#define DART_HD __host__ __device__ __forceinline__
#define DART_UNROLL _Pragma("unroll")
#define DART_PRAGMA_(x) _Pragma(#x)
#define DART_UNROLL_N(k) DART_PRAGMA_(unroll k)
#else
#define DART_HD inline
#define DART_UNROLL
#define DART_UNROLL_N(k)
#endif

#ifdef DART_PHASE_CLOCK
#include <cstdio>
#ifndef DART_PHASE_CLOCK_BLOCK
#define DART_PHASE_CLOCK_BLOCK (blockIdx.x == 0)
#endif
#ifdef __CUDA_ARCH__
#define DART_CLOCK() clock64()
#else
#define DART_CLOCK() 0LL
#endif
#endif

#ifndef DART_TREE_SUMS
#define DART_TREE_SUMS 1
#endif
// vector recursions of the tiled sweeps (forward_tile, corrector_tile): the lanes hand the new vector to each other through
// the shared-memory copy they store anyway (1) instead of 2 n SHFL per stage (0)
#ifndef DART_XCHG_SMEM
#define DART_XCHG_SMEM 1
#endif
#ifndef DART_SWEEP_UNROLL
#define DART_SWEEP_UNROLL 3
#endif

namespace dart {

enum Status : int32_t { ST_CONVERGED = 0, ST_MAXITER = 1, ST_INFEASIBLE = 2, ST_NUMERIC = 3, ST_ACCEPTABLE = 4 };
// severity order used when the axes of one instance are combined (acceptable ranks just above converged)
DART_HD int status_rank(int32_t st) { return st == ST_CONVERGED ? 0 : st == ST_ACCEPTABLE ? 1 : st == ST_MAXITER ? 2 : st == ST_INFEASIBLE ? 3 : 4; }

struct SolverOpts {
    double tol, mu0, kappa_mu, theta_mu, kappa_eps, tau_min, bound_push, eta, smax;
    double s_phi, s_theta, delta_sw, gamma_theta, gamma_phi, theta_small;
    int max_iter, max_backtrack;
    // IPOPT's acceptable-level termination: acc_iter consecutive iterates with error <= acc_tol (acc_iter = 0: off)
    double acc_tol;
    int acc_iter;
    // 1: Mehrotra predictor-corrector barrier strategy where the kernel has it (the scan path of the PMPC axis problems)
    int mehrotra;
    int cold;          // this launch carries no warm plan (set per launch by dart_solve from warm_w == NULL)
    int mu0_auto;      // mu_init was left at its default: predictor-corrector solves start from Model::MU0_PC (Solver::start_mu)
};

DART_HD double dmax(double a, double b) { return a > b ? a : b; }
DART_HD double dmin(double a, double b) { return a < b ? a : b; }

// two consecutive doubles at a 16-byte aligned address: one 128-bit shared-memory access on the device
struct alignas(16) D2 { double x, y; };
DART_HD D2 ld2(const double* p) {
#ifdef __CUDA_ARCH__
    return *reinterpret_cast<const D2*>(p);
#else
    D2 r; r.x = p[0]; r.y = p[1]; return r;
#endif
}
DART_HD void st2(double* p, double x, double y) {
#ifdef __CUDA_ARCH__
    D2 v; v.x = x; v.y = y;
    *reinterpret_cast<D2*>(p) = v;
#else
    p[0] = x; p[1] = y;
#endif
}
// load c consecutive doubles (c compile-time) from an aligned address with 128-bit accesses (+ one 64-bit tail)
template <int C>
DART_HD void ldv(const double* p, double* out) {
    DART_UNROLL for (int i = 0; i + 1 < C; i += 2) { const D2 v = ld2(p + i); out[i] = v.x; out[i + 1] = v.y; }
    if (C & 1) out[C - 1] = p[C - 1];
}

// Per-tile result / hand-shake slot (doubles) at the head of each problem's workspace stride.
constexpr int kSlot = 4;

// Where the problems of this thread block live in shared memory: problem q's slot is base + q*stride, its workspace
// follows the slot.  Used by the serial-sweep phase, in which thread `tid` runs the sweeps of problem `tid`.
struct BlockCtx {
    double* base;
    int stride, nprob, tid;
};

// Host stand-in for a tile of one lane (and a block of one problem).
struct HostTile {
    static constexpr int kLanes = 1;
    DART_HD int lane() const { return 0; }
    DART_HD int size() const { return 1; }
    DART_HD void sync() const {}
    DART_HD void block_sync() const {}
    DART_HD bool block_any(bool p) const { return p; }
    DART_HD bool warp_any(bool p) const { return p; }
    static constexpr bool kLockstep = false;
    DART_HD unsigned warp_ballot(bool p) const { return p ? 1u : 0u; }
    DART_HD bool group_all(bool p) const { return p; }
    DART_HD double shfl(double v, int) const { return v; }
    DART_HD double sum(double v) const { return v; }
    DART_HD double max(double v) const { return v; }
    DART_HD double max_nonneg(double v) const { return v; }
    DART_HD double min_pos(double v) const { return v; }
    DART_HD double min(double v) const { return v; }
};

// Every lane runs the whole loop redundantly (same values, same addresses): no exchange, no syncs.
struct SerialTile {
    static constexpr int kLanes = 1;
    DART_HD int lane() const { return 0; }
    DART_HD int size() const { return 1; }
    DART_HD void sync() const {}
};

// ------------------------------------------------------------------------------------------------------
// Parallel-in-time LQ sweeps for 2-state / 1-input stage problems (the PMPC axis problem).
//
// The Riccati recursion is a fold of "conditional value functions" under an associative combination rule (Sarkka &
// Garcia-Fernandez, "Temporal parallelization of dynamic programming and linear quadratic control", 2021), so the
// value functions of ALL stages follow from a suffix scan: log2(N+1) combination levels instead of N dependent stages.
// An element describes an interval i -> j of the horizon,
//     V_{i->j}(x_i, x_j) = max_lam [ 1/2 x_i' J x_i + q' x_i - 1/2 lam' C lam - lam' (x_j - A x_i - b) ],
// one stage k -> k+1 with diagonal-free coupling (H_xu = 0) is  A = A_k, b = d_k - B g_u / H_uu, C = B B' / H_uu,
// J = H_xx, q = g_x, and the terminal cost is the element A = 0, b = 0, C = 0, J = P_N, q = p_N.  The combination of
// stage k with everything after it has J = P_k, q = p_k (the value function of the serial recursion).
namespace scan2 {
struct El {
    double a00, a01, a10, a11, b0, b1, c00, c01, c11, q0, q1, j00, j01, j11;
};
// O = L (+) R: L covers the earlier interval, R the later one.
DART_HD void combine(const El& L, const El& R, El& O) {
    const double t00 = 1.0 + L.c00 * R.j00 + L.c01 * R.j01, t01 = L.c00 * R.j01 + L.c01 * R.j11;
    const double t10 = L.c01 * R.j00 + L.c11 * R.j01, t11 = 1.0 + L.c01 * R.j01 + L.c11 * R.j11;
    const double id = 1.0 / (t00 * t11 - t01 * t10);               // det(I + C_L J_R) >= 1 for positive semidefinite C, J
    const double m00 = t11 * id, m01 = -t01 * id, m10 = -t10 * id, m11 = t00 * id;
    const double am00 = R.a00 * m00 + R.a01 * m10, am01 = R.a00 * m01 + R.a01 * m11;
    const double am10 = R.a10 * m00 + R.a11 * m10, am11 = R.a10 * m01 + R.a11 * m11;
    const double v0 = L.b0 - (L.c00 * R.q0 + L.c01 * R.q1), v1 = L.b1 - (L.c01 * R.q0 + L.c11 * R.q1);
    const double x00 = am00 * L.c00 + am01 * L.c01, x01 = am00 * L.c01 + am01 * L.c11;
    const double x10 = am10 * L.c00 + am11 * L.c01, x11 = am10 * L.c01 + am11 * L.c11;
    const double r0 = R.q0 + R.j00 * L.b0 + R.j01 * L.b1, r1 = R.q1 + R.j01 * L.b0 + R.j11 * L.b1;
    const double s0 = m00 * r0 + m10 * r1, s1 = m01 * r0 + m11 * r1;
    const double y00 = m00 * R.j00 + m10 * R.j01, y01 = m00 * R.j01 + m10 * R.j11;
    const double y10 = m01 * R.j00 + m11 * R.j01, y11 = m01 * R.j01 + m11 * R.j11;
    const double z00 = L.a00 * y00 + L.a10 * y10, z01 = L.a00 * y01 + L.a10 * y11;
    const double z10 = L.a01 * y00 + L.a11 * y10, z11 = L.a01 * y01 + L.a11 * y11;
    El o;
    o.a00 = am00 * L.a00 + am01 * L.a10; o.a01 = am00 * L.a01 + am01 * L.a11;
    o.a10 = am10 * L.a00 + am11 * L.a10; o.a11 = am10 * L.a01 + am11 * L.a11;
    o.b0 = am00 * v0 + am01 * v1 + R.b0; o.b1 = am10 * v0 + am11 * v1 + R.b1;
    o.c00 = x00 * R.a00 + x01 * R.a01 + R.c00;
    o.c01 = 0.5 * ((x00 * R.a10 + x01 * R.a11) + (x10 * R.a00 + x11 * R.a01)) + R.c01;
    o.c11 = x10 * R.a10 + x11 * R.a11 + R.c11;
    o.q0 = L.a00 * s0 + L.a10 * s1 + L.q0; o.q1 = L.a01 * s0 + L.a11 * s1 + L.q1;
    o.j00 = z00 * L.a00 + z01 * L.a10 + L.j00;
    o.j01 = 0.5 * ((z00 * L.a01 + z01 * L.a11) + (z10 * L.a00 + z11 * L.a10)) + L.j01;
    o.j11 = z10 * L.a01 + z11 * L.a11 + L.j11;
    O = o;
}
// J and q of L (+) R only (the value function of L's first stage; what the last combination of a scan needs)
DART_HD void combine_jq(const El& L, const El& R, double& j00, double& j01, double& j11, double& q0, double& q1) {
    const double t00 = 1.0 + L.c00 * R.j00 + L.c01 * R.j01, t01 = L.c00 * R.j01 + L.c01 * R.j11;
    const double t10 = L.c01 * R.j00 + L.c11 * R.j01, t11 = 1.0 + L.c01 * R.j01 + L.c11 * R.j11;
    const double id = 1.0 / (t00 * t11 - t01 * t10);
    const double m00 = t11 * id, m01 = -t01 * id, m10 = -t10 * id, m11 = t00 * id;
    const double r0 = R.q0 + R.j00 * L.b0 + R.j01 * L.b1, r1 = R.q1 + R.j01 * L.b0 + R.j11 * L.b1;
    const double s0 = m00 * r0 + m10 * r1, s1 = m01 * r0 + m11 * r1;
    const double y00 = m00 * R.j00 + m10 * R.j01, y01 = m00 * R.j01 + m10 * R.j11;
    const double y10 = m01 * R.j00 + m11 * R.j01, y11 = m01 * R.j01 + m11 * R.j11;
    const double z00 = L.a00 * y00 + L.a10 * y10, z01 = L.a00 * y01 + L.a10 * y11;
    const double z10 = L.a01 * y00 + L.a11 * y10, z11 = L.a01 * y01 + L.a11 * y11;
    q0 = L.a00 * s0 + L.a10 * s1 + L.q0; q1 = L.a01 * s0 + L.a11 * s1 + L.q1;
    j00 = z00 * L.a00 + z01 * L.a10 + L.j00;
    j01 = 0.5 * ((z00 * L.a01 + z01 * L.a11) + (z10 * L.a00 + z11 * L.a10)) + L.j01;
    j11 = z10 * L.a01 + z11 * L.a11 + L.j11;
}
// affine maps x -> M x + v of the closed loop: O = later (after) earlier
struct Af {
    double m00, m01, m10, m11, v0, v1;
};
DART_HD void compose(const Af& later, const Af& earlier, Af& O) {
    Af o;
    o.m00 = later.m00 * earlier.m00 + later.m01 * earlier.m10; o.m01 = later.m00 * earlier.m01 + later.m01 * earlier.m11;
    o.m10 = later.m10 * earlier.m00 + later.m11 * earlier.m10; o.m11 = later.m10 * earlier.m01 + later.m11 * earlier.m11;
    o.v0 = later.m00 * earlier.v0 + later.m01 * earlier.v1 + later.v0;
    o.v1 = later.m10 * earlier.v0 + later.m11 * earlier.v1 + later.v1;
    O = o;
}
}  // namespace scan2

template <class M>
struct Workspace {
    static constexpr int n = M::NX, m = M::NU, nr = M::NR, ny = n + m, nc = n + m + 1, np = M::NP;
    // A, Bm: physical NP x NP / NP x NU blocks only (carried-input rows are structural); PP, HS: packed symmetric
    static constexpr int npa = np * np, npb = np * m, nps = n * (n + 1) / 2;
    // stage Hessian: only its structural non-zeros are stored -- the diagonal and, for models with carried inputs,
    // the (carried input j, input j) pairs that the tilt-rate cost and the tilt-rate rows couple
    static constexpr int NH = ny + (M::NAUG > 0 ? M::NAUG : 0);
    DART_HD static constexpr int hslot(int i, int c) {
        const int lo = i < c ? i : c, hi = i < c ? c : i;
        if (lo == hi) return lo;
        if (M::NAUG > 0 && lo >= np && lo < n && hi == n + (lo - np)) return ny + (lo - np);
        return -1;
    }
    DART_HD static constexpr bool rows_fit() {
        for (int r = 0; r < nr; ++r)
            if (M::row_ib(r) >= 0 && hslot(M::row_ia(r), M::row_ib(r)) < 0) return false;
        return true;
    }
    static_assert(rows_fit(), "a two-variable constraint row couples variables outside the stored Hessian pattern");
    // kVec (the models whose Riccati sweep runs across the tile): the arrays the sweeps gather from are laid out so that
    // what one lane reads per stage is CONTIGUOUS and 16-byte aligned -- A and B column-major (a lane of the backward sweep
    // multiplies one column of [A B d]), even strides, every array at an even offset -- so a column is two 128-bit loads
    // at immediate offsets from one per-lane pointer instead of four 64-bit gathers with computed addresses.
    static constexpr bool kVec = !M::SERIAL_RICCATI;
    DART_HD static constexpr int ev(int x) { return (x + 1) & ~1; }
    // even stride >= x that is an ODD multiple of 16 bytes: lanes working on consecutive stages then cover the banks
    // with 128-bit accesses (and are at worst 2-way conflicted with 64-bit ones)
    DART_HD static constexpr int pad2(int x) { return (ev(x) % 4 == 2) ? ev(x) : ev(x) + 2; }
    // per-stage strides (in doubles) of the arrays that lanes index by stage.  Scalar-access arrays: odd, so that lanes
    // working on consecutive stages hit distinct shared-memory banks (16 banks of 8 bytes)
    static constexpr int sA = kVec ? pad2(npa) : (npa | 1), sB = kVec ? pad2(npb) : (npb | 1), sD = kVec ? pad2(n) : n;
    static constexpr int sH = NH | 1, sG = ny | 1, sK = kVec ? ev(m * n) : ((m * n) | 1);
    static constexpr int sM = ny | 1;
    // index of dF_a/dx_b, dF_a/du_j inside a stage block (column-major for kVec)
    DART_HD static constexpr int aidx(int a, int b) { return kVec ? b * np + a : a * np + b; }
    DART_HD static constexpr int bidx(int a, int j) { return kVec ? j * np + a : a * m + j; }
    // stage scratch MM of the tiled Riccati sweep: the full P_{k+1} (row stride rP), W = P [A B d] column-major over the
    // nq = np + m variables with computed curvature + the d column (column stride rP), then the packed upper triangle
    // of the stage matrix (+ gradient column)
    static constexpr int nq = np + m, rP = ev(n), nPF = n * rP, nW = (nq + 1) * rP, nMq = nq * (nq + 1) / 2 + nq;
    static constexpr int nscr = M::SERIAL_RICCATI ? sM * (ny + 1) : nPF + nW + ev(nMq);
    double *X, *U, *A, *Bm, *D, *LAM, *PP, *PV, *K, *KFF, *DX, *DU, *BL;
    double *S, *ZL, *ZU, *ISL, *ISU, *RC, *DS, *MM, *TANU, *HS, *GR, *ZR, *REF;

    // every array starts at an even offset (16-byte aligned when the workspace is)
    DART_HD static int doubles(int N) {
        return ev((N + 1) * n) + ev(N * m) + ev(N * sA) + ev(N * sB) + ev(N * sD) + ev(N * n) + ev((N + 1) * nps) + ev((N + 1) * n) +
               ev(N * sK) + ev(N * m) + ev((N + 1) * n) + ev(N * m) + ev(N * m) + 7 * ev(N * nr) + ev(nscr) + ev(N * m) +
               ev(N * sH) + ev(N * sG) + 2 + ev(M::ref_doubles(N));
    }
    DART_HD void bind(double* p, int N) {
        X = p;   p += ev((N + 1) * n);
        U = p;   p += ev(N * m);
        A = p;   p += ev(N * sA);
        Bm = p;  p += ev(N * sB);
        D = p;   p += ev(N * sD);
        LAM = p; p += ev(N * n);
        PP = p;  p += ev((N + 1) * nps);
        PV = p;  p += ev((N + 1) * n);
        K = p;   p += ev(N * sK);
        KFF = p; p += ev(N * m);
        DX = p;  p += ev((N + 1) * n);
        DU = p;  p += ev(N * m);
        BL = p;  p += ev(N * m);
        S = p;   p += ev(N * nr);
        ZL = p;  p += ev(N * nr);
        ZU = p;  p += ev(N * nr);
        ISL = p; p += ev(N * nr);
        ISU = p; p += ev(N * nr);
        RC = p;  p += ev(N * nr);
        DS = p;  p += ev(N * nr);
        MM = p;  p += ev(nscr);
        TANU = p; p += ev(N * m);
        HS = p;  p += ev(N * sH);
        GR = p;  p += ev(N * sG);
        ZR = p;  p += 2;          // a stored 0.0: gather target for structural zeros
        REF = p;
    }
};

// ------------------------------------------------------------------------------------------------------
// NC > 0: horizon known at compile time (constant offsets, unrollable stage loops); NC == 0: runtime horizon.
template <class M, class T, int NC = 0>
struct Solver {
    static constexpr int n = M::NX, m = M::NU, nr = M::NR, ny = n + m, nc = n + m + 1, np = M::NP;
    using Prm = typename M::Prm;
    using W = Workspace<M>;

    const T& tile;
    const Prm& prm;
    const SolverOpts& o;
    const int N;
    W& w;
    const BlockCtx& bc;

    DART_HD Solver(const T& t, const Prm& p, const SolverOpts& oo, int NN, W& ww, const BlockCtx& b)
        : tile(t), prm(p), o(oo), N(NC > 0 ? NC : NN), w(ww), bc(b) {}

    static constexpr int npa = W::npa, npb = W::npb, nps = W::nps;
    static constexpr int sA = W::sA, sB = W::sB, sD = W::sD, sH = W::sH, sG = W::sG, sK = W::sK, sM = W::sM;
    static constexpr bool kVec = W::kVec;
    // packed upper-triangular index of a symmetric d x d matrix
    DART_HD static constexpr int sidx(int i, int j, int d) {
        return i <= j ? i * d - i * (i - 1) / 2 + (j - i) : j * d - j * (j - 1) / 2 + (i - j);
    }
    // full stage Jacobians from the stored physical blocks (carried-input rows: A = 0, B = identity)
    DART_HD double Aat(int k, int a, int b) const { return (a < np && b < np) ? w.A[k * sA + W::aidx(a, b)] : 0.0; }
    DART_HD double Bat(int k, int a, int j) const { return (a < np) ? w.Bm[k * sB + W::bidx(a, j)] : ((a - np) == j ? 1.0 : 0.0); }
    DART_HD double Pat(int k, int a, int b) const { return w.PP[k * nps + sidx(a, b, n)]; }
    DART_HD double Hat(int k, int i, int c) const { return W::hslot(i, c) < 0 ? 0.0 : w.HS[k * sH + (W::hslot(i, c) < 0 ? 0 : W::hslot(i, c))]; }
    // gather descriptor of H[i][c] (structural zeros read the stored 0.0)
    DART_HD void hgat(int i, int c, int& off, int& ks) const {
        const int sl = W::hslot(i, c);
        if (sl < 0) { off = (int)(w.ZR - w.X); ks = 0; } else { off = (int)(w.HS - w.X) + sl; ks = sH; }
    }

    DART_HD double yval(int k, int i) const { return i < n ? w.X[k * n + i] : w.U[k * m + (i - n)]; }
    DART_HD double rowval(int k, int r) const {
        double t = M::row_sa(r) * yval(k, M::row_ia(r));
        if (M::row_ib(r) >= 0) t += M::row_sb(r) * yval(k, M::row_ib(r));
        return t;
    }
    DART_HD static bool masked(int k, int r) { return k == 0 && M::row_skip0(r); }

    // cost gradient wrt y_i at stage k (stage cost incl. tilt-rate term)
    DART_HD double cost_grad(int k, int i) const {
        double g = 2.0 * M::wy(prm, i) * (yval(k, i) - M::ry(prm, w.REF, k, i));
        if (M::NAUG > 0) {
            if (i >= n) {
                int j = i - n;
                g += 2.0 * M::wd(prm, j) * (w.U[k * m + j] - w.X[k * n + np + j]);
            } else if (i >= np) {
                int j = i - np;
                g -= 2.0 * M::wd(prm, j) * (w.U[k * m + j] - w.X[k * n + np + j]);
            }
        }
        return g;
    }

    // reciprocals of nr positive numbers from ONE division (prefix products); returns their product
    DART_HD static double batch_inv(const double* v, double* inv) {
        double pre[nr];
        double tot = 1.0;
        DART_UNROLL for (int r = 0; r < nr; ++r) { pre[r] = tot; tot = (r == 0) ? v[0] : tot * v[r]; }
        double t = 1.0 / tot;
        DART_UNROLL for (int r = nr - 1; r >= 0; --r) {
            inv[r] = (r == 0) ? t : t * pre[r];
            if (r > 0) t *= v[r];
        }
        return tot;
    }

    // ---- E1: dynamics, Jacobians, defects, objective, barrier logs, constraint violation (stage-parallel)
    DART_HD void eval1(double& f, double& L, double& th, double& pinf) {
        double f_ = 0.0, L_ = 0.0, th_ = 0.0, pi_ = 0.0;
        for (int k = tile.lane(); k < N; k += tile.size()) {
            double x[n], u[m], F[n], Aloc[np * np], Bloc[np * m];
            DART_UNROLL for (int i = 0; i < n; ++i) x[i] = w.X[k * n + i];
            DART_UNROLL for (int j = 0; j < m; ++j) u[j] = w.U[k * m + j];
            double tanu[m];
            M::dyn(prm, x, u, F, Aloc, Bloc, tanu);
            DART_UNROLL for (int j = 0; j < m; ++j) w.TANU[k * m + j] = tanu[j];
            if (kVec) {
                // column-major stage blocks, written as 128-bit pairs (np is even for these models)
                DART_UNROLL for (int b = 0; b < np; ++b)
                    DART_UNROLL for (int a = 0; a + 1 < np; a += 2) st2(&w.A[k * sA + b * np + a], Aloc[a * np + b], Aloc[(a + 1) * np + b]);
                DART_UNROLL for (int j = 0; j < m; ++j)
                    DART_UNROLL for (int a = 0; a + 1 < np; a += 2) st2(&w.Bm[k * sB + j * np + a], Bloc[a * m + j], Bloc[(a + 1) * m + j]);
            } else {
                DART_UNROLL for (int i = 0; i < npa; ++i) w.A[k * sA + i] = Aloc[i];
                DART_UNROLL for (int i = 0; i < npb; ++i) w.Bm[k * sB + i] = Bloc[i];
            }
            DART_UNROLL for (int a = np; a < n; ++a) F[a] = u[a - np];
            DART_UNROLL for (int a = 0; a < n; ++a) {
                double d = F[a] - w.X[(k + 1) * n + a];
                w.D[k * sD + a] = d;
                th_ += fabs(d);
                pi_ = dmax(pi_, fabs(d));
            }
            DART_UNROLL for (int i = 0; i < ny; ++i) {
                double e = (i < n ? x[i] : u[i - n]) - M::ry(prm, w.REF, k, i);
                f_ += M::wy(prm, i) * e * e;
            }
            if (M::NAUG > 0) {
                DART_UNROLL for (int j = 0; j < m; ++j) {
                    double e = u[j] - x[np + j];
                    f_ += M::wd(prm, j) * e * e;
                }
            }
            // rows: one reciprocal and one logarithm per STAGE (of the product of the rows' slack products: at most
            // nr <= 6 factors >= 1e-24, no underflow), the per-row reciprocals follow from prefix products
            double sl[nr], su[nr], prod[nr];
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                double lo, hi;
                M::bounds(prm, r, lo, hi);
                const double s = w.S[k * nr + r];
                sl[r] = s - lo; su[r] = hi - s;
                if (masked(k, r)) { prod[r] = 1.0; continue; }
                const double rc = rowval(k, r) - s;
                prod[r] = sl[r] * su[r];
                w.RC[k * nr + r] = rc;
                th_ += fabs(rc);
                pi_ = dmax(pi_, fabs(rc));
            }
            double ip[nr];
            const double ptot = batch_inv(prod, ip);
            L_ += log(ptot);
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) continue;
                w.ISL[k * nr + r] = su[r] * ip[r];
                w.ISU[k * nr + r] = sl[r] * ip[r];
            }
        }
        if (tile.lane() == 0) {
            DART_UNROLL for (int i = 0; i < n; ++i) {
                double e = w.X[N * n + i] - M::rT(prm, w.REF, N, i);
                f_ += M::wT(prm, i) * e * e;
            }
        }
        f = tile.sum(f_);
        L = tile.sum(L_);
        th = tile.sum(th_);
        pinf = tile.max_nonneg(pi_);
        tile.sync();
    }

    // ---- E2: KKT residuals with the current multipliers (stage-parallel)
    DART_HD void eval2(double& dinf, double& zs_min, double& zs_max, double& lam_sum, double& z_sum) {
        double di = 0.0, zmn = 1e300, zmx = 0.0, ls = 0.0, zs = 0.0;
        for (int k = tile.lane(); k < N; k += tile.size()) {
            double lam[n];
            DART_UNROLL for (int a = 0; a < n; ++a) { lam[a] = w.LAM[k * n + a]; ls += fabs(lam[a]); }
            double g[ny];
            DART_UNROLL for (int i = 0; i < ny; ++i) g[i] = cost_grad(k, i);
            DART_UNROLL for (int i = 0; i < n; ++i) {
                double acc = 0.0;
                DART_UNROLL for (int a = 0; a < np; ++a)
                    if (i < np) acc += w.A[k * sA + W::aidx(a, i)] * lam[a];
                g[i] += acc - (k >= 1 ? w.LAM[(k - 1) * n + i] : 0.0);
            }
            DART_UNROLL for (int j = 0; j < m; ++j) {
                double acc = 0.0;
                DART_UNROLL for (int a = 0; a < np; ++a) acc += w.Bm[k * sB + W::bidx(a, j)] * lam[a];
                if (M::NAUG > 0) acc += lam[np + j];
                w.BL[k * m + j] = acc;
                g[n + j] += acc;
            }
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) continue;
                double zl = w.ZL[k * nr + r], zu = w.ZU[k * nr + r];
                double nu = zu - zl;
                g[M::row_ia(r)] += M::row_sa(r) * nu;
                if (M::row_ib(r) >= 0) g[M::row_ib(r)] += M::row_sb(r) * nu;
                double lo, hi;
                M::bounds(prm, r, lo, hi);
                double s = w.S[k * nr + r];
                double a = zl * (s - lo), b = zu * (hi - s);
                zmn = dmin(zmn, dmin(a, b));
                zmx = dmax(zmx, dmax(a, b));
                zs += fabs(zl) + fabs(zu);
            }
            DART_UNROLL for (int i = 0; i < ny; ++i)
                if (i >= n || k >= 1) di = dmax(di, fabs(g[i]));
        }
        if (tile.lane() == 0) {
            DART_UNROLL for (int i = 0; i < n; ++i) {
                double gT = 2.0 * M::wT(prm, i) * (w.X[N * n + i] - M::rT(prm, w.REF, N, i)) - w.LAM[(N - 1) * n + i];
                di = dmax(di, fabs(gT));
            }
        }
        dinf = tile.max_nonneg(di);
        zs_min = tile.min_pos(zmn);
        zs_max = tile.max_nonneg(zmx);
        lam_sum = tile.sum(ls);
        z_sum = tile.sum(zs);
        tile.sync();
    }

    // m x m SPD "factorisation" + solve.  m = 1: reciprocal; m = 2: explicit inverse from the determinant (one
    // division, no square root); larger m: Cholesky.  Returns false when the block is not positive definite.
    DART_HD static bool chol(double* H) {
        if (m == 1) {
            if (!(H[0] > 0.0)) return false;
            H[0] = 1.0 / H[0];
            return true;
        }
        if (m == 2) {
            const double a = H[0], b = 0.5 * (H[1] + H[2]), d = H[3];
            const double det = a * d - b * b;
            if (!(a > 0.0) || !(det > 0.0)) return false;
            const double id = 1.0 / det;
            H[0] = d * id; H[1] = -b * id; H[2] = -b * id; H[3] = a * id;
            return true;
        }
        DART_UNROLL for (int j = 0; j < m; ++j) {
            double d = H[j * m + j];
            DART_UNROLL for (int q = 0; q < j; ++q) d -= H[j * m + q] * H[j * m + q];
            if (!(d > 0.0)) return false;
            d = sqrt(d);
            H[j * m + j] = d;
            DART_UNROLL for (int i = j + 1; i < m; ++i) {
                double v = H[i * m + j];
                DART_UNROLL for (int q = 0; q < j; ++q) v -= H[i * m + q] * H[j * m + q];
                H[i * m + j] = v / d;
            }
        }
        return true;
    }
    DART_HD static void chol_solve(const double* Lc, double* b) {
        if (m == 1) { b[0] *= Lc[0]; return; }
        if (m == 2) {
            const double b0 = b[0], b1 = b[1];
            b[0] = Lc[0] * b0 + Lc[1] * b1;
            b[1] = Lc[2] * b0 + Lc[3] * b1;
            return;
        }
        DART_UNROLL for (int i = 0; i < m; ++i) {
            double v = b[i];
            DART_UNROLL for (int q = 0; q < i; ++q) v -= Lc[i * m + q] * b[q];
            b[i] = v / Lc[i * m + i];
        }
        DART_UNROLL for (int i = m - 1; i >= 0; --i) {
            double v = b[i];
            DART_UNROLL for (int q = i + 1; q < m; ++q) v -= Lc[q * m + i] * b[q];
            b[i] = v / Lc[i * m + i];
        }
    }

    // ---- stage Hessian blocks and condensed gradients for the current mu and multipliers (stage-parallel)
    DART_HD void prep(double mu) {
        for (int k = tile.lane(); k < N; k += tile.size()) {
            double H[ny * ny], g[ny];
            DART_UNROLL for (int i = 0; i < ny * ny; ++i) H[i] = 0.0;
            DART_UNROLL for (int i = 0; i < ny; ++i) { H[i * ny + i] = 2.0 * M::wy(prm, i); g[i] = cost_grad(k, i); }
            if (M::NAUG > 0) {
                DART_UNROLL for (int j = 0; j < m; ++j) {
                    const double wd2 = 2.0 * M::wd(prm, j);
                    H[(n + j) * ny + n + j] += wd2;
                    H[(np + j) * ny + np + j] += wd2;
                    H[(n + j) * ny + np + j] -= wd2;
                    H[(np + j) * ny + n + j] -= wd2;
                }
            }
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) continue;
                const int ia = M::row_ia(r), ib = M::row_ib(r);
                const double sa = M::row_sa(r), sb = M::row_sb(r);
                const double isl = w.ISL[k * nr + r], isu = w.ISU[k * nr + r];
                const double sig = w.ZL[k * nr + r] * isl + w.ZU[k * nr + r] * isu;
                const double nuhat = mu * (isu - isl) + sig * w.RC[k * nr + r];
                H[ia * ny + ia] += sig * sa * sa;
                g[ia] += sa * nuhat;
                if (ib >= 0) {
                    H[ib * ny + ib] += sig * sb * sb;
                    H[ia * ny + ib] += sig * sa * sb;
                    H[ib * ny + ia] += sig * sa * sb;
                    g[ib] += sb * nuhat;
                }
            }
            // Lagrangian curvature of the tilt input: -tan(u_j) (B^T lambda)_j
            DART_UNROLL for (int j = 0; j < m; ++j) H[(n + j) * ny + n + j] += -w.TANU[k * m + j] * w.BL[k * m + j];
            DART_UNROLL for (int i = 0; i < ny; ++i)
                DART_UNROLL for (int c = i; c < ny; ++c)
                    if (W::hslot(i, c) >= 0) w.HS[k * sH + W::hslot(i, c)] = H[i * ny + c];
            DART_UNROLL for (int i = 0; i < ny; ++i) w.GR[k * sG + i] = g[i];
        }
        if (M::SERIAL_RICCATI && tile.lane() == 0) {
            // terminal value function P_N, p_N for the serial sweep (which runs without the model parameters)
            DART_UNROLL for (int i = 0; i < n; ++i) {
                DART_UNROLL for (int j = i; j < n; ++j) w.PP[N * nps + sidx(i, j, n)] = (i == j) ? 2.0 * M::wT(prm, i) : 0.0;
                w.PV[N * n + i] = 2.0 * M::wT(prm, i) * (w.X[N * n + i] - M::rT(prm, w.REF, N, i));
            }
        }
        tile.sync();
    }

    // ---- Riccati backward sweep, one matrix ELEMENT per lane.  Each stage is three rounds separated by tile syncs:
    //   1. W = P_{k+1} [A B d] + [0 0 p_{k+1}]            n x (nq+1) elements
    //   2. M = [H g] + [A B]^T W  (upper triangle + g)     nq(nq+1)/2 + nq elements
    //   3. pivot inverse, gains K, P_k = Mxx + Mxu K       n(n+1)/2 + n elements
    // with nq = np + m: rows/columns of carried-input states are structural (A = 0, B = I), so their entries of M are
    // those of [H g] and are read from there.  What a lane reads depends only on its element, not on the stage, and it
    // is contiguous: round 1 multiplies one ROW of the full P_{k+1} (kept in a scratch block next to the packed history)
    // with one COLUMN of [A B d] (stored column-major), round 2 one column of [A B] with one column of W (written
    // column-major by round 1).  So a lane holds one pointer per operand -- decoded once, before the stage loop, stepped
    // by the stage stride -- and the loop body is 128-bit loads at immediate offsets, a short FMA chain and one store per
    // round (measured before this layout: 357 instructions per stage of which 58 were FP64 arithmetic and 93 IMADs).
    struct Gat {
        int off, ks;                                    // element of stage k: w.X[off + k * ks]
    };
    DART_HD int offs(const double* q) const { return (int)(q - w.X); }
    DART_HD static void tri_decode(int e, int d, int& i, int& c) {   // inverse of sidx(i, c, d) for i <= c
        i = 0;
        while (e >= d - i) { e -= d - i; ++i; }
        c = i + e;
    }
    template <class TL>
    DART_HD void backward(const TL& tl) {
        constexpr int G = TL::kLanes;
        constexpr int nq = W::nq, ncw = nq + 1, ntri = nq * (nq + 1) / 2, rP = W::rP;
        constexpr int nW = n * ncw, nM = W::nMq, nPt = n * (n + 1) / 2, nP = nPt + n;
        constexpr int PW = (nW + G - 1) / G, PM = (nM + G - 1) / G, PQ = (nP + G - 1) / G;
        constexpr int NZ = (M::NAUG > 0) ? M::NAUG : 1;
        static_assert(np % 2 == 0, "the vectorised sweep reads columns of [A B d] as 128-bit pairs");
        const int lane = tl.lane();
        double* const base = w.X;
        double* const PF = w.MM;                        // full P_{k+1}, row stride rP
        double* const WW = w.MM + W::nPF;               // W column-major, column stride rP
        double* const MQ = WW + W::nW;
        const int oA = offs(w.A), oB = offs(w.Bm), oD = offs(w.D), oG = offs(w.GR), oMQ = offs(MQ);
        const int oPP = offs(w.PP), oPV = offs(w.PV), oK = offs(w.K), oKF = offs(w.KFF);

        // terminal value function: packed history entry, full scratch copy, p_N
        for (int c = lane; c <= n; c += G) {
            if (c < n) {
                DART_UNROLL for (int i = 0; i < n; ++i) {
                    const double v = (i == c) ? 2.0 * M::wT(prm, c) : 0.0;
                    if (i <= c) w.PP[N * nps + sidx(i, c, n)] = v;
                    PF[i * rP + c] = v;
                }
            } else {
                DART_UNROLL for (int i = 0; i < n; ++i)
                    w.PV[N * n + i] = 2.0 * M::wT(prm, i) * (w.X[N * n + i] - M::rT(prm, w.REF, N, i));
            }
        }

        // ---- round 1: element (a, cw); cw < np: state column, cw < nq: input column, cw == nq: d/p column
        struct DW { const double* prow; const double* tcol; double* wout; int tKs, pv; double gsel, ez[NZ]; bool act; };
        DW dw[PW];
        DART_UNROLL for (int ps = 0; ps < PW; ++ps) {
            DW& d = dw[ps];
            const int e = lane + ps * G;
            d.act = e < nW;
            const int ee = d.act ? e : 0;
            const int cw = ee / n, a = ee % n;
            d.prow = PF + a * rP;
            d.pv = oPV + n + a;
            d.gsel = (cw == nq) ? 1.0 : 0.0;
            DART_UNROLL for (int z = 0; z < NZ; ++z) d.ez[z] = (M::NAUG > 0 && cw >= np && cw < nq && cw - np == z) ? 1.0 : 0.0;
            if (cw < np) { d.tcol = w.A + (N - 1) * sA + cw * np; d.tKs = sA; }
            else if (cw < nq) { d.tcol = w.Bm + (N - 1) * sB + (cw - np) * np; d.tKs = sB; }
            else { d.tcol = w.D + (N - 1) * sD; d.tKs = sD; }
            d.wout = WW + cw * rP + a;
        }
        // ---- round 2: element (iq, cq) of the upper triangle over the nq active variables, or (iq, g)
        struct DM { const double* tcol; const double* wcol; int tKs, hOff, hKs; double augc[NZ]; bool act; };
        DM dm[PM];
        DART_UNROLL for (int ps = 0; ps < PM; ++ps) {
            DM& d = dm[ps];
            const int e = lane + ps * G;
            d.act = e < nM;
            const int ee = d.act ? e : 0;
            int iq, cq;
            if (ee < ntri) tri_decode(ee, nq, iq, cq); else { iq = ee - ntri; cq = nq; }
            const int fi = iq < np ? iq : n + (iq - np);              // index among the ny stage variables
            if (iq < np) { d.tcol = w.A + (N - 1) * sA + iq * np; d.tKs = sA; }
            else { d.tcol = w.Bm + (N - 1) * sB + (iq - np) * np; d.tKs = sB; }
            d.wcol = WW + cq * rP;
            if (cq < nq) { const int fc = cq < np ? cq : n + (cq - np); hgat(fi, fc, d.hOff, d.hKs); }
            else { d.hOff = oG + fi; d.hKs = sG; }
            // input row: + W[np + j][.] (B = I on the carried inputs)
            DART_UNROLL for (int z = 0; z < NZ; ++z) d.augc[z] = (M::NAUG > 0 && iq >= np && iq - np == z) ? 1.0 : 0.0;
        }
        // ---- round 3: element (i, c) of the upper triangle of P_k, or p_k[i]
        struct DQ { Gat mu[m], mic, miu[m], st; int kOff, kJs, kKs, pf0, pf1; bool storeK, act, isP; };
        DQ dq[PQ];
        DART_UNROLL for (int ps = 0; ps < PQ; ++ps) {
            DQ& d = dq[ps];
            const int e = lane + ps * G;
            d.act = e < nP;
            const int ee = d.act ? e : 0;
            int i, c;
            if (ee < nPt) tri_decode(ee, n, i, c); else { i = ee - nPt; c = n; }
            DART_UNROLL for (int j = 0; j < m; ++j) {
                // M[u_j][c]
                if (c == n) d.mu[j] = Gat{oMQ + ntri + np + j, 0};
                else if (c < np) d.mu[j] = Gat{oMQ + sidx(c, np + j, nq), 0};
                else hgat(c, n + j, d.mu[j].off, d.mu[j].ks);
                // M[i][u_j]
                if (i < np) d.miu[j] = Gat{oMQ + sidx(i, np + j, nq), 0};
                else hgat(i, n + j, d.miu[j].off, d.miu[j].ks);
            }
            if (c == n) d.mic = (i < np) ? Gat{oMQ + ntri + i, 0} : Gat{oG + i, sG};
            else if (i < np && c < np) d.mic = Gat{oMQ + sidx(i, c, nq), 0};
            else hgat(i, c, d.mic.off, d.mic.ks);
            d.st = (c == n) ? Gat{oPV + i, n} : Gat{oPP + sidx(i, c, n), nps};
            d.isP = d.act && c < n;
            d.pf0 = (c < n) ? i * rP + c : 0;
            d.pf1 = (c < n) ? c * rP + i : 0;
            d.storeK = d.act && i == 0;
            if (c == n) { d.kOff = oKF; d.kJs = 1; d.kKs = m; }
            else { d.kOff = oK + c; d.kJs = n; d.kKs = sK; }
        }
        tl.sync();

        for (int k = N - 1; k >= 0; --k) {
            // carried-input defects of this stage (the d column of the structural rows), one uniform address
            double dz[NZ];
            if (M::NAUG > 0) { DART_UNROLL for (int z = 0; z < M::NAUG; ++z) dz[z] = w.D[k * sD + np + z]; }
            DART_UNROLL for (int ps = 0; ps < PW; ++ps) {
                DW& d = dw[ps];
                double pr[n], tc[np];
                ldv<n>(d.prow, pr);
                ldv<np>(d.tcol, tc);
                d.tcol -= d.tKs;
                // three short independent chains instead of one of depth np + NAUG + 1 (the sweep is latency-bound)
                double s0 = pr[0] * tc[0], s1 = pr[np / 2] * tc[np / 2], s2 = d.gsel * base[d.pv + k * n];
                DART_UNROLL for (int t = 1; t < np / 2; ++t) s0 += pr[t] * tc[t];
                DART_UNROLL for (int t = np / 2 + 1; t < np; ++t) s1 += pr[t] * tc[t];
                if (M::NAUG > 0) {
                    DART_UNROLL for (int z = 0; z < M::NAUG; ++z) s2 += pr[np + z] * (d.ez[z] + d.gsel * dz[z]);
                }
#if DART_TREE_SUMS
                if (d.act) *d.wout = (s0 + s1) + s2;
#else
                {
                    double acc = d.gsel * base[d.pv + k * n];
                    DART_UNROLL for (int t = 0; t < np; ++t) acc += pr[t] * tc[t];
                    if (M::NAUG > 0) { DART_UNROLL for (int z = 0; z < M::NAUG; ++z) acc += pr[np + z] * (d.ez[z] + d.gsel * dz[z]); }
                    if (d.act) *d.wout = acc;
                }
#endif
            }
            tl.sync();
            DART_UNROLL for (int ps = 0; ps < PM; ++ps) {
                DM& d = dm[ps];
                double tc[np], wc[n];
                ldv<np>(d.tcol, tc);
                d.tcol -= d.tKs;
                ldv<n>(d.wcol, wc);
                double s0 = base[d.hOff + k * d.hKs], s1 = tc[np / 2] * wc[np / 2];
                DART_UNROLL for (int a = 0; a < np / 2; ++a) s0 += tc[a] * wc[a];
                DART_UNROLL for (int a = np / 2 + 1; a < np; ++a) s1 += tc[a] * wc[a];
                if (M::NAUG > 0) { DART_UNROLL for (int z = 0; z < M::NAUG; ++z) s1 += d.augc[z] * wc[np + z]; }
#if DART_TREE_SUMS
                if (d.act) MQ[lane + ps * G] = s0 + s1;
#else
                {
                    double acc = base[d.hOff + k * d.hKs];
                    DART_UNROLL for (int a = 0; a < np; ++a) acc += tc[a] * wc[a];
                    if (M::NAUG > 0) { DART_UNROLL for (int z = 0; z < M::NAUG; ++z) acc += d.augc[z] * wc[np + z]; }
                    if (d.act) MQ[lane + ps * G] = acc;
                }
#endif
            }
            tl.sync();
            // every lane inverts the same m x m pivot block (+ escalating shift if it is not positive definite)
            double Lc[m * m];
            DART_UNROLL for (int i = 0; i < m; ++i)
                DART_UNROLL for (int j = 0; j < m; ++j) Lc[i * m + j] = MQ[sidx(np + i, np + j, nq)];
            if (!chol(Lc)) {
                // cold path (kept rolled: it is almost never taken and must not bloat the stage loop)
                double shift = 1e-4;
                DART_UNROLL_N(1)
                for (int tries = 0; tries < 40; ++tries) {
                    DART_UNROLL for (int i = 0; i < m; ++i)
                        DART_UNROLL for (int j = 0; j < m; ++j)
                            Lc[i * m + j] = MQ[sidx(np + i, np + j, nq)] + (i == j ? shift : 0.0);
                    if (chol(Lc)) break;
                    shift *= 8.0;
                }
            }
            DART_UNROLL for (int ps = 0; ps < PQ; ++ps) {
                const DQ& d = dq[ps];
                double kt[m];
                DART_UNROLL for (int j = 0; j < m; ++j) kt[j] = -base[d.mu[j].off + k * d.mu[j].ks];
                chol_solve(Lc, kt);
                if (d.storeK) { DART_UNROLL for (int j = 0; j < m; ++j) base[d.kOff + j * d.kJs + k * d.kKs] = kt[j]; }
                if (k > 0) {                         // P_0 / p_0 are never used (x_0 is fixed)
                    double v = base[d.mic.off + k * d.mic.ks];
                    DART_UNROLL for (int j = 0; j < m; ++j) v += base[d.miu[j].off + k * d.miu[j].ks] * kt[j];
                    if (d.act) base[d.st.off + k * d.st.ks] = v;
                    if (d.isP) { PF[d.pf0] = v; PF[d.pf1] = v; }
                }
            }
            tl.sync();
        }
    }

    // ---- Riccati backward sweep for small problems, run by ONE thread per problem (the serial-sweep phase of run()):
    // P, p and the stage matrix stay in REGISTERS, so the loop-carried dependency never touches shared memory, and
    // stage data loads do not depend on P, so they are issued ahead of the dependent chain.  Needs only the workspace
    // (terminal P_N, p_N were written by prep()).
    DART_HD void backward_serial() {
        // structure known at compile time (M::a_kind: 0 general, 1 exact zero, 2 exact one; the stage Hessian's
        // pattern W::hslot) drops loads and FMAs from the single-thread chain; M is formed as its upper triangle
        auto tkind = [](int b, int c) { return c < n ? M::a_kind(b, c) : 0; };          // entry (b, c) of [A B d]
        auto hzero = [](int i, int c) { return c < ny && W::hslot(i, c) < 0; };
        double P[n * n], pv[n];
        DART_UNROLL for (int i = 0; i < n; ++i) {
            DART_UNROLL for (int j = 0; j < n; ++j) P[i * n + j] = Pat(N, i, j);
            pv[i] = w.PV[N * n + i];
        }
        double Tm[n * nc], HG[ny * nc];
        auto load = [&](int k, double* T_, double* H_) {
            DART_UNROLL for (int a = 0; a < n; ++a) {
                DART_UNROLL for (int b = 0; b < n; ++b) T_[a * nc + b] = (tkind(a, b) == 0) ? Aat(k, a, b) : (tkind(a, b) == 2 ? 1.0 : 0.0);
                DART_UNROLL for (int j = 0; j < m; ++j) T_[a * nc + n + j] = Bat(k, a, j);
                T_[a * nc + ny] = w.D[k * sD + a];
            }
            DART_UNROLL for (int i = 0; i < ny; ++i) {
                DART_UNROLL for (int c = i; c < ny; ++c) H_[i * nc + c] = hzero(i, c) ? 0.0 : Hat(k, i, c);
                H_[i * nc + ny] = w.GR[k * sG + i];
            }
        };
        load(N - 1, Tm, HG);
        DART_UNROLL_N(DART_SWEEP_UNROLL)
        for (int k = N - 1; k >= 0; --k) {
            // next stage's data first: these loads do not depend on P, keep them ahead of the dependent chain
            double Tn[n * nc], Hn[ny * nc];
            load(k > 0 ? k - 1 : 0, Tn, Hn);
            double Wm[n * nc], Mm[ny * nc];
            DART_UNROLL for (int a = 0; a < n; ++a)
                DART_UNROLL for (int c = 0; c < nc; ++c) {
                    double acc = 0.0;
                    bool first = true;
                    if (c == nc - 1) { acc = pv[a]; first = false; }
                    DART_UNROLL for (int b = 0; b < n; ++b) {
                        const int kd = tkind(b, c);
                        if (kd == 1) continue;
                        const double term_p = P[a * n + b];
                        if (first) { acc = (kd == 2) ? term_p : term_p * Tm[b * nc + c]; first = false; }
                        else if (kd == 2) acc += term_p;
                        else acc += term_p * Tm[b * nc + c];
                    }
                    Wm[a * nc + c] = acc;
                }
            DART_UNROLL for (int i = 0; i < ny; ++i)
                DART_UNROLL for (int c = 0; c < nc; ++c) {
                    if (c < ny && c < i) continue;                     // upper triangle (+ gradient column) only
                    double acc = 0.0;
                    bool first = hzero(i, c);
                    if (!first) acc = HG[i * nc + c];
                    DART_UNROLL for (int a = 0; a < n; ++a) {
                        const int kd = (i < n) ? tkind(a, i) : 0;
                        if (kd == 1) continue;
                        const double wv = Wm[a * nc + c];
                        if (first) { acc = (kd == 2) ? wv : Tm[a * nc + i] * wv; first = false; }
                        else if (kd == 2) acc += wv;
                        else acc += Tm[a * nc + i] * wv;
                    }
                    Mm[i * nc + c] = acc;
                }
            auto Mu = [&](int i, int c) { return (c < ny && c < i) ? Mm[c * nc + i] : Mm[i * nc + c]; };
            double Lc[m * m];
            if (m == 1) {
                // scalar pivot: reciprocal on the fast path, escalating shift only if it is not positive
                double h = Mm[n * nc + n];
                if (!(h > 0.0)) {
                    double shift = 1e-4;
                    while (!(h + shift > 0.0) && shift < 1e30) shift *= 8.0;
                    h += shift;
                }
                Lc[0] = 1.0 / h;
            } else {
                double shift = 0.0;
                for (int tries = 0; tries < 40; ++tries) {
                    DART_UNROLL for (int i = 0; i < m; ++i)
                        DART_UNROLL for (int j = 0; j < m; ++j) Lc[i * m + j] = Mu(n + i, n + j) + (i == j ? shift : 0.0);
                    if (chol(Lc)) break;
                    shift = (shift == 0.0) ? 1e-4 : shift * 8.0;
                }
            }
            double Kt[m * (n + 1)];      // columns 0..n-1: feedback gains, column n: feed-forward
            DART_UNROLL for (int c = 0; c <= n; ++c) {
                double kt[m];
                const int cc = (c < n) ? c : ny;
                DART_UNROLL for (int j = 0; j < m; ++j) kt[j] = -Mu(n + j, cc);
                chol_solve(Lc, kt);
                DART_UNROLL for (int j = 0; j < m; ++j) Kt[j * (n + 1) + c] = kt[j];
            }
            if (k > 0) {
                DART_UNROLL for (int i = 0; i < n; ++i) {
                    DART_UNROLL for (int c = i; c <= n; ++c) {
                        const int cc = (c < n) ? c : ny;
                        double v = Mm[i * nc + cc];
                        DART_UNROLL for (int j = 0; j < m; ++j) v += Mm[i * nc + n + j] * Kt[j * (n + 1) + c];
                        if (c < n) { P[i * n + c] = v; P[c * n + i] = v; } else pv[i] = v;
                    }
                }
            }
            DART_UNROLL for (int j = 0; j < m; ++j) {
                DART_UNROLL for (int c = 0; c < n; ++c) w.K[k * sK + j * n + c] = Kt[j * (n + 1) + c];
                w.KFF[k * m + j] = Kt[j * (n + 1) + n];
            }
            if (k > 0) {
                DART_UNROLL for (int i = 0; i < n; ++i)
                    DART_UNROLL for (int j = i; j < n; ++j) w.PP[k * nps + sidx(i, j, n)] = P[i * n + j];
                DART_UNROLL for (int i = 0; i < n; ++i) w.PV[k * n + i] = pv[i];
            }
            DART_UNROLL for (int i = 0; i < n * nc; ++i) Tm[i] = Tn[i];
            DART_UNROLL for (int i = 0; i < ny * nc; ++i) HG[i] = Hn[i];
        }
    }

    // ---- forward sweep, ONE thread per problem: the short affine recurrence runs in registers.  The stage data of
    // step k+1 is loaded before the dependent arithmetic of step k so shared-memory latency stays off the chain.
    DART_HD void forward() {
        double dx[n];
        DART_UNROLL for (int i = 0; i < n; ++i) { dx[i] = 0.0; w.DX[i] = 0.0; }
        double Kc[m * n], kc[m], Ac[n * n], Bc[n * m], dc[n];
        auto load = [&](int k, double* K_, double* k_, double* A_, double* B_, double* d_) {
            DART_UNROLL for (int i = 0; i < m * n; ++i) K_[i] = w.K[k * sK + i];
            DART_UNROLL for (int i = 0; i < m; ++i) k_[i] = w.KFF[k * m + i];
            DART_UNROLL for (int a = 0; a < n; ++a) {
                DART_UNROLL for (int b = 0; b < n; ++b) A_[a * n + b] = (M::a_kind(a, b) == 0) ? Aat(k, a, b) : 0.0;
                DART_UNROLL for (int j = 0; j < m; ++j) B_[a * m + j] = Bat(k, a, j);
            }
            DART_UNROLL for (int i = 0; i < n; ++i) d_[i] = w.D[k * sD + i];
        };
        load(0, Kc, kc, Ac, Bc, dc);
        DART_UNROLL_N(DART_SWEEP_UNROLL)
        for (int k = 0; k < N; ++k) {
            double Kn[m * n], kn[m], An[n * n], Bn[n * m], dn[n];
            const int kk = (k + 1 < N) ? k + 1 : k;
            load(kk, Kn, kn, An, Bn, dn);
            double du[m], nx[n];
            DART_UNROLL for (int j = 0; j < m; ++j) {
                double acc = kc[j];
                DART_UNROLL for (int i = 0; i < n; ++i) acc += Kc[j * n + i] * dx[i];
                du[j] = acc;
            }
            DART_UNROLL for (int a = 0; a < n; ++a) {
                double acc = dc[a];
                DART_UNROLL for (int i = 0; i < n; ++i) {
                    if (M::a_kind(a, i) == 1) continue;               // structural zero / one of the sensitivity
                    if (M::a_kind(a, i) == 2) acc += dx[i]; else acc += Ac[a * n + i] * dx[i];
                }
                DART_UNROLL for (int j = 0; j < m; ++j) acc += Bc[a * m + j] * du[j];
                nx[a] = acc;
            }
            DART_UNROLL for (int j = 0; j < m; ++j) w.DU[k * m + j] = du[j];
            DART_UNROLL for (int a = 0; a < n; ++a) w.DX[(k + 1) * n + a] = nx[a];
            DART_UNROLL for (int a = 0; a < n; ++a) dx[a] = nx[a];
            DART_UNROLL for (int i = 0; i < m * n; ++i) Kc[i] = Kn[i];
            DART_UNROLL for (int i = 0; i < m; ++i) kc[i] = kn[i];
            DART_UNROLL for (int i = 0; i < n * n; ++i) Ac[i] = An[i];
            DART_UNROLL for (int i = 0; i < n * m; ++i) Bc[i] = Bn[i];
            DART_UNROLL for (int i = 0; i < n; ++i) dc[i] = dn[i];
        }
    }

    // ---- forward sweep across the tile (larger models): every lane keeps dx_k in registers and forms du_k; lane a
    // forms row a of dx_{k+1} = A dx + B du + d, and the rows are exchanged by shuffles -- per stage one n-deep and one
    // m-deep FMA chain plus a shuffle instead of the whole stage in one lane.  Needs at least n lanes.  (Measured: forming
    // the closed-loop row A + B K off the chain instead costs 12 more FMAs per stage and is 3 % slower.)
    template <class TL>
    DART_HD void forward_tile(const TL& tl) {
        const int lane = tl.lane();
        const int a = lane < n ? lane : n - 1;               // lanes >= n shadow the last row (no stores)
        const int ap = a < np ? a : np - 1;                  // row of the stored physical blocks this lane reads
        const bool phys = a < np;
        double sel[m];                                       // carried-input rows: x_{k+1}[np + j] = u_k[j] (+ defect)
        DART_UNROLL for (int j = 0; j < m; ++j) sel[j] = (M::NAUG > 0 && a - np == j) ? 1.0 : 0.0;
        double dx[n];
        DART_UNROLL for (int i = 0; i < n; ++i) dx[i] = 0.0;
        w.DX[a] = 0.0;
        // stage data is loaded one stage ahead: it does not depend on dx, only the FMA chain and the shuffle do
        double Kc[m * n], kc[m], Ar[np], Br[m], dc;
        auto load = [&](int k, double* K_, double* k_, double* A_, double* B_, double& d_) {
            if (kVec) ldv<m * n>(&w.K[k * sK], K_);
            else { DART_UNROLL for (int i = 0; i < m * n; ++i) K_[i] = w.K[k * sK + i]; }
            DART_UNROLL for (int j = 0; j < m; ++j) k_[j] = w.KFF[k * m + j];
            DART_UNROLL for (int i = 0; i < np; ++i) A_[i] = w.A[k * sA + W::aidx(ap, i)];
            DART_UNROLL for (int j = 0; j < m; ++j) B_[j] = w.Bm[k * sB + W::bidx(ap, j)];
            d_ = w.D[k * sD + a];
        };
        load(0, Kc, kc, Ar, Br, dc);
        for (int k = 0; k < N; ++k) {
            double Kn[m * n], kn[m], An[np], Bn[m], dn;
            load(k + 1 < N ? k + 1 : k, Kn, kn, An, Bn, dn);
            double du[m];
            DART_UNROLL for (int j = 0; j < m; ++j) {
                double acc = kc[j];
                DART_UNROLL for (int i = 0; i < n; ++i) acc += Kc[j * n + i] * dx[i];
                du[j] = acc;
            }
            double vp = dc, va = dc;
            DART_UNROLL for (int i = 0; i < np; ++i) vp += Ar[i] * dx[i];
            DART_UNROLL for (int j = 0; j < m; ++j) { vp += Br[j] * du[j]; va += sel[j] * du[j]; }
            const double v = (M::NAUG > 0 && !phys) ? va : vp;
            // unconditional stores: lanes >= n shadow row n - 1 and every lane holds du, so the extra lanes write the same
            // values to the same addresses (one wavefront) -- no divergent branch per stage (it held 4.4 % of RMPC's samples)
            w.DX[(k + 1) * n + a] = v;
            DART_UNROLL for (int j = 0; j < m; ++j) w.DU[k * m + j] = du[j];
#if DART_XCHG_SMEM
            tl.sync();
            if (n % 2 == 0) ldv<n>(&w.DX[(k + 1) * n], dx);
            else { DART_UNROLL for (int i = 0; i < n; ++i) dx[i] = w.DX[(k + 1) * n + i]; }
#else
            DART_UNROLL for (int i = 0; i < n; ++i) dx[i] = tl.shfl(v, i);
#endif
            DART_UNROLL for (int i = 0; i < m * n; ++i) Kc[i] = Kn[i];
            DART_UNROLL for (int j = 0; j < m; ++j) { kc[j] = kn[j]; Br[j] = Bn[j]; }
            DART_UNROLL for (int i = 0; i < np; ++i) Ar[i] = An[i];
            dc = dn;
        }
    }

    // ---- Mehrotra predictor-corrector step for the models whose Riccati sweep runs across the tile (the scan path has its
    // own: sweeps_scan_pc).  The predictor is the ordinary sweep pair on the affine-scaling right-hand side (prep(0)).
    // pc_rows(): the predictor's slack / multiplier steps, its longest steps, mu = sigma * mean(z s) with
    // sigma = (mean after the longest affine step / mean)^3 (oracle/ipm.py, Options.mehrotra), and the CHANGE of the stage
    // gradients that the corrector's complementarity target (s + ds)(z + dz) = mu - ds_a dz_a makes against the
    // predictor's; it goes into GR (free after the backward sweep), the predictor's row steps into DS (post() reads them).
    DART_HD double pc_rows(double mu_min, double inv_pairs) {
        double rp = 0.0, rd = 0.0, comp = 0.0;
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) { w.DS[k * nr + r] = 0.0; continue; }
                const int ia = M::row_ia(r), ib = M::row_ib(r);
                double dy = M::row_sa(r) * ((ia < n) ? w.DX[k * n + ia] : w.DU[k * m + (ia - n)]);
                if (ib >= 0) dy += M::row_sb(r) * ((ib < n) ? w.DX[k * n + ib] : w.DU[k * m + (ib - n)]);
                const double dsa = dy + w.RC[k * nr + r];
                const double isl = w.ISL[k * nr + r], isu = w.ISU[k * nr + r];
                double lo, hi;
                M::bounds(prm, r, lo, hi);
                const double sv = w.S[k * nr + r];
                w.DS[k * nr + r] = dsa;
                rp = dmax(rp, dmax(-dsa * isl, dsa * isu));
                rd = dmax(rd, dmax(1.0 + isl * dsa, 1.0 - isu * dsa));          // -dz_l / z_l, -dz_u / z_u of the predictor
                comp += w.ZL[k * nr + r] * (sv - lo) + w.ZU[k * nr + r] * (hi - sv);
            }
        }
        rp = tile.max_nonneg(rp);
        rd = tile.max_nonneg(rd);
        comp = tile.sum(comp);
        const double apa = (rp > 1.0) ? 1.0 / rp : 1.0, ada = (rd > 1.0) ? 1.0 / rd : 1.0;
        double caff = 0.0;
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) continue;
                const double dsa = w.DS[k * nr + r];
                const double zl = w.ZL[k * nr + r], zu = w.ZU[k * nr + r];
                double lo, hi;
                M::bounds(prm, r, lo, hi);
                const double sv = w.S[k * nr + r];
                const double dzl = -zl - zl * w.ISL[k * nr + r] * dsa, dzu = -zu + zu * w.ISU[k * nr + r] * dsa;
                caff += ((sv - lo) + apa * dsa) * (zl + ada * dzl) + ((hi - sv) - apa * dsa) * (zu + ada * dzu);
            }
        }
        caff = tile.sum(caff);
        const double ratio = caff / comp;
        const double sigma = dmin(1.0, dmax(1e-8, ratio * ratio * ratio));
        const double mu = dmax(mu_min, sigma * comp * inv_pairs);
        for (int k = tile.lane(); k < N; k += tile.size()) {
            double g[ny];
            DART_UNROLL for (int i = 0; i < ny; ++i) g[i] = 0.0;
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) continue;
                const double dsa = w.DS[k * nr + r];
                const double isl = w.ISL[k * nr + r], isu = w.ISU[k * nr + r];
                const double zl = w.ZL[k * nr + r], zu = w.ZU[k * nr + r];
                const double cl = dsa * (-zl - zl * isl * dsa), cu = dsa * (-zu + zu * isu * dsa);
                const double dg = (mu + cu) * isu - (mu - cl) * isl;
                g[M::row_ia(r)] += M::row_sa(r) * dg;
                if (M::row_ib(r) >= 0) g[M::row_ib(r) >= 0 ? M::row_ib(r) : 0] += M::row_sb(r) * dg;
            }
            DART_UNROLL for (int i = 0; i < ny; ++i) w.GR[k * sG + i] = g[i];
        }
        tile.sync();
        return mu;
    }

    // The corrector's solve with the predictor's factorisation.  Only the stage gradients differ (dg, in GR), so with the
    // gains K_k, the inverse pivots and the stage Jacobians the corrected direction costs one backward VECTOR recursion,
    //       dm_u = dg_u + B' dp_{k+1},   dp_k = dg_x + A' dp_{k+1} + K' dm_u,   dkff_k = -Muu^-1 dm_u,
    // and then the ordinary forward sweep again with the feed-forward terms kff + dkff (it is linear in them: the result is
    // the corrected step itself, with plain stores -- no second code path, no read-modify-write on the sweep's chain).
    // Lane a forms row a of dp_k (every lane forms the m entries of dm_u), rows are exchanged by shuffles as in the
    // forward sweep; dp_k and dm_u go to DX / DU, which are free once pc_rows() has used the predictor's step.
    // The inverse pivots are not stored (the RMPC workspace has no room: 8 problems per SM): the stage matrix's block
    // M[u][carried inputs] is the stored diagonal h_j = H[u_j][x_{np+j}] (tilt-rate cost and rate rows; never zero, the rate
    // rows' barrier terms are in it), and K[:, np+j] = -Muu^-1 e_j h_j, so  -Muu^-1 dm_u = sum_j K[:, np+j] dm_u[j] / h_j.
    // feedforward() does that and p_k += dp_k stage-parallel (the divisions run once, not once per stage of the recursion).
    DART_HD void feedforward() {
        static_assert(M::NAUG == m, "the inverse pivot is recovered from the gains of the carried-input columns");
        for (int k = tile.lane(); k < N; k += tile.size()) {
            double sc[m];
            DART_UNROLL for (int j = 0; j < m; ++j) sc[j] = w.DU[k * m + j] / w.HS[k * sH + W::hslot(np + j, n + j)];
            DART_UNROLL for (int i = 0; i < m; ++i) {
                double acc = w.KFF[k * m + i];
                DART_UNROLL for (int j = 0; j < m; ++j) acc += w.K[k * sK + i * n + np + j] * sc[j];
                w.KFF[k * m + i] = acc;
            }
            if (k > 0) { DART_UNROLL for (int i = 0; i < n; ++i) w.PV[k * n + i] += w.DX[k * n + i]; }
        }
        tile.sync();
    }
    template <class TL>
    DART_HD void corrector_tile(const TL& tl) {
        static_assert(kVec && np % 2 == 0, "corrector_tile: tiled-sweep models");
        const int lane = tl.lane();
        const int a = lane < n ? lane : n - 1;
        const bool phys = a < np;
        const int ac = phys ? a : np - 1;
        double dp[n];
        DART_UNROLL for (int i = 0; i < n; ++i) dp[i] = 0.0;
        for (int k = N - 1; k >= 0; --k) {
            double Ac[np], Bc[np * m], Kc[m], mug[m];
            ldv<np>(&w.A[k * sA + ac * np], Ac);                 // column ac of A_k: dF_b / dx_ac
            ldv<np * m>(&w.Bm[k * sB], Bc);                      // column-major: dF_b / du_j at [j * np + b]
            DART_UNROLL for (int j = 0; j < m; ++j) Kc[j] = w.K[k * sK + j * n + a];
            double t = w.GR[k * sG + a];
            DART_UNROLL for (int j = 0; j < m; ++j) {
                // two short chains per entry
                double s0 = w.GR[k * sG + n + j] + dp[np + j];   // carried-input rows: B = identity
                double s1 = Bc[j * np + np / 2] * dp[np / 2];
                DART_UNROLL for (int b = 0; b < np / 2; ++b) s0 += Bc[j * np + b] * dp[b];
                DART_UNROLL for (int b = np / 2 + 1; b < np; ++b) s1 += Bc[j * np + b] * dp[b];
                mug[j] = s0 + s1;
            }
            if (phys) { DART_UNROLL for (int b = 0; b < np; ++b) t += Ac[b] * dp[b]; }
            double v = t;
            DART_UNROLL for (int j = 0; j < m; ++j) v += Kc[j] * mug[j];
            DART_UNROLL for (int j = 0; j < m; ++j) w.DU[k * m + j] = mug[j];      // same value from every lane (see forward_tile)
            w.DX[k * n + a] = v;
#if DART_XCHG_SMEM
            tl.sync();
            if (n % 2 == 0) ldv<n>(&w.DX[k * n], dp);
            else { DART_UNROLL for (int i = 0; i < n; ++i) dp[i] = w.DX[k * n + i]; }
#else
            DART_UNROLL for (int i = 0; i < n; ++i) dp[i] = tl.shfl(v, i);
#endif
        }
        tl.sync();
    }

    // the same recursion by one lane in plain loops (tiles narrower than the state, and the host build's 1-lane tile)
    DART_HD void corrector_serial_backward() {
        double dp[n];
        DART_UNROLL for (int i = 0; i < n; ++i) dp[i] = 0.0;
        for (int k = N - 1; k >= 0; --k) {
            double mug[m], v[n];
            for (int j = 0; j < m; ++j) {
                double acc = w.GR[k * sG + n + j];
                for (int b = 0; b < n; ++b) acc += Bat(k, b, j) * dp[b];
                mug[j] = acc;
            }
            for (int i = 0; i < n; ++i) {
                double acc = w.GR[k * sG + i];
                for (int b = 0; b < n; ++b) acc += Aat(k, b, i) * dp[b];
                for (int j = 0; j < m; ++j) acc += w.K[k * sK + j * n + i] * mug[j];
                v[i] = acc;
            }
            for (int j = 0; j < m; ++j) w.DU[k * m + j] = mug[j];
            for (int i = 0; i < n; ++i) { w.DX[k * n + i] = v[i]; dp[i] = v[i]; }
        }
    }

    // ---- both sweeps of the Newton step as SCANS across the tile (2-state / 1-input problems, one lane per stage plus one
    // for the terminal cost: lanes == N + 1).  Backward: suffix scan of the conditional value functions (scan2::combine,
    // log2(N+1) = 4 levels instead of N = 15 dependent Riccati stages) gives P_k, p_k of every stage at once; the gains
    // are then stage-parallel; forward: prefix scan of the closed-loop affine maps.  Everything stays in registers and
    // shuffles; no other tile is involved, so tiles of a block run and finish independently (no block barriers).
    // Pivot regularisation: the serial sweep shifts the pivot H_uu + B'P B when it is not positive; here the shift is
    // decided on H_uu alone (it must be inverted before P is known).  B'P B >= 0, so the pivot is then positive too;
    // the two rules differ only when H_uu <= 0 < H_uu + B'P B, and only in that iteration's step, not in the fixed point.
    template <class TL>
    DART_HD void sweeps_scan(const TL& tl) {
        using scan2::El;
        using scan2::Af;
        constexpr int G = TL::kLanes;
        const int lane = tl.lane();
        const bool stage = lane < N;
        const int k = stage ? lane : N - 1;                 // the terminal lane reads stage N-1 (values unused)
        const double a00 = Aat(k, 0, 0), a01 = Aat(k, 0, 1), a10 = Aat(k, 1, 0), a11 = Aat(k, 1, 1);
        const double B0 = w.Bm[k * sB + 0], B1 = w.Bm[k * sB + 1];
        const double d0 = w.D[k * sD + 0], d1 = w.D[k * sD + 1];
        const double gu = w.GR[k * sG + 2];
        double huu = w.HS[k * sH + 2];
        if (!(huu > 0.0)) {
            double shift = 1e-4;
            while (!(huu + shift > 0.0) && shift < 1e30) shift *= 8.0;
            huu += shift;
        }
        const double ih = 1.0 / huu;
        El e;
        if (stage) {
            e.a00 = a00; e.a01 = a01; e.a10 = a10; e.a11 = a11;
            const double t = ih * gu;
            e.b0 = d0 - B0 * t; e.b1 = d1 - B1 * t;
            e.c00 = B0 * B0 * ih; e.c01 = B0 * B1 * ih; e.c11 = B1 * B1 * ih;
            e.q0 = w.GR[k * sG + 0]; e.q1 = w.GR[k * sG + 1];
            e.j00 = w.HS[k * sH + 0]; e.j01 = 0.0; e.j11 = w.HS[k * sH + 1];
        } else {
            e.a00 = e.a01 = e.a10 = e.a11 = e.b0 = e.b1 = e.c00 = e.c01 = e.c11 = 0.0;
            e.j00 = w.PP[N * nps + 0]; e.j01 = w.PP[N * nps + 1]; e.j11 = w.PP[N * nps + 2];
            e.q0 = w.PV[N * n + 0]; e.q1 = w.PV[N * n + 1];
        }
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            const int src = (lane + off) & (G - 1);
            El r;
            r.a00 = tl.shfl(e.a00, src); r.a01 = tl.shfl(e.a01, src); r.a10 = tl.shfl(e.a10, src); r.a11 = tl.shfl(e.a11, src);
            r.b0 = tl.shfl(e.b0, src); r.b1 = tl.shfl(e.b1, src);
            r.c00 = tl.shfl(e.c00, src); r.c01 = tl.shfl(e.c01, src); r.c11 = tl.shfl(e.c11, src);
            r.q0 = tl.shfl(e.q0, src); r.q1 = tl.shfl(e.q1, src);
            r.j00 = tl.shfl(e.j00, src); r.j01 = tl.shfl(e.j01, src); r.j11 = tl.shfl(e.j11, src);
            if (lane + off <= N) {
                if (2 * off >= G) scan2::combine_jq(e, r, e.j00, e.j01, e.j11, e.q0, e.q1);   // last level: only P, p are still needed
                else scan2::combine(e, r, e);
            }
        }
        // value function of the next stage: lane k needs P_{k+1}, p_{k+1}
        const int nxt = (lane + 1) & (G - 1);
        const double P00 = tl.shfl(e.j00, nxt), P01 = tl.shfl(e.j01, nxt), P11 = tl.shfl(e.j11, nxt);
        const double pv0 = tl.shfl(e.q0, nxt), pv1 = tl.shfl(e.q1, nxt);
        const double PB0 = P00 * B0 + P01 * B1, PB1 = P01 * B0 + P11 * B1;
        const double ihh = 1.0 / (huu + B0 * PB0 + B1 * PB1);
        const double Pd0 = P00 * d0 + P01 * d1 + pv0, Pd1 = P01 * d0 + P11 * d1 + pv1;
        const double K0 = -ihh * (PB0 * a00 + PB1 * a10), K1 = -ihh * (PB0 * a01 + PB1 * a11);
        const double kff = -ihh * (gu + B0 * Pd0 + B1 * Pd1);
        Af f;
        if (stage) {
            f.m00 = a00 + B0 * K0; f.m01 = a01 + B0 * K1; f.m10 = a10 + B1 * K0; f.m11 = a11 + B1 * K1;
            f.v0 = B0 * kff + d0; f.v1 = B1 * kff + d1;
        } else {
            f.m00 = 1.0; f.m01 = 0.0; f.m10 = 0.0; f.m11 = 1.0; f.v0 = 0.0; f.v1 = 0.0;
        }
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            const int src = (lane - off) & (G - 1);
            Af r;
            r.m00 = tl.shfl(f.m00, src); r.m01 = tl.shfl(f.m01, src); r.m10 = tl.shfl(f.m10, src); r.m11 = tl.shfl(f.m11, src);
            r.v0 = tl.shfl(f.v0, src); r.v1 = tl.shfl(f.v1, src);
            if (lane >= off) scan2::compose(f, r, f);
        }
        // dx_{k+1} = f.v of lane k; dx_k comes from lane k-1 (dx_0 = 0)
        const int prv = (lane - 1) & (G - 1);
        double x0 = tl.shfl(f.v0, prv), x1 = tl.shfl(f.v1, prv);
        if (lane == 0) { x0 = 0.0; x1 = 0.0; w.DX[0] = 0.0; w.DX[1] = 0.0; }
        if (stage) {
            w.DU[k] = K0 * x0 + K1 * x1 + kff;
            w.DX[(k + 1) * n + 0] = f.v0; w.DX[(k + 1) * n + 1] = f.v1;
            if (k + 1 < N) {                                 // P_N, p_N are already in place (prep)
                w.PP[(k + 1) * nps + 0] = P00; w.PP[(k + 1) * nps + 1] = P01; w.PP[(k + 1) * nps + 2] = P11;
                w.PV[(k + 1) * n + 0] = pv0; w.PV[(k + 1) * n + 1] = pv1;
            }
        }
        tl.sync();
    }

    // ---- Mehrotra predictor-corrector step on the scan path (one lane per stage; nr == 1: the input bound row).
    // The slow instances of a batch hold a weakly active bound: slack and multiplier of that pair must shrink together, and
    // a Newton step on z s = mu only quarters the product (both factors halve), so each barrier reduction of the monotone
    // schedule cost them 3-5 iterations (tools/ipm_variants.py: 9.1 -> 5.7 iterations on average, 16 -> 10 at most).
    //   predictor: the affine-scaling direction (prep(0) has built the gradient with mu = 0) by the scans below;
    //   barrier parameter: mu = sigma * mean(z s), sigma = (mean(z s) after the longest affine step / mean(z s))^3;
    //   corrector: (s + ds)(z + dz) = mu keeps the predictor's product ds dz.  Against the predictor's right-hand side only
    //   the input gradients change, so the correction is solved for separately with the SAME factorisation: the vector
    //   parts of both sweeps are affine recursions in the closed-loop maps Phi_k = A_k + B_k K_k,
    //       dp_k = Phi_k' dp_{k+1} + K_k' dg_k,   dkff_k = -(dg_k + B_k' dp_{k+1}) / h_k,   ddx_{k+1} = Phi_k ddx_k + B_k dkff_k,
    //   i.e. two more scans of 2x2 affine maps (6 values per element instead of 14, no inversions).
    // Writes DX, DU, PP, PV of the corrected direction, the product terms for post() (CL, CU) and returns the new mu.
    DART_HD double* CL() const { return w.KFF; }      // ds dz_l of the predictor per row (the serial sweeps' gain arrays are free here)
    DART_HD double* CU() const { return w.K; }
    template <class TL>
    DART_HD double sweeps_scan_pc(const TL& tl, double mu_min) {
        static_assert(nr == 1 && m == 1 && n == 2, "predictor-corrector step: 2-state / 1-input problems with one bound row");
        using scan2::El;
        using scan2::Af;
        constexpr int G = TL::kLanes;
        const int lane = tl.lane();
        const bool stage = lane < N;
        const int k = stage ? lane : N - 1;
        const double a00 = Aat(k, 0, 0), a01 = Aat(k, 0, 1), a10 = Aat(k, 1, 0), a11 = Aat(k, 1, 1);
        const double B0 = w.Bm[k * sB + 0], B1 = w.Bm[k * sB + 1];
        const double d0 = w.D[k * sD + 0], d1 = w.D[k * sD + 1];
        const double gu = w.GR[k * sG + 2];
        double huu = w.HS[k * sH + 2];
        if (!(huu > 0.0)) {
            double shift = 1e-4;
            while (!(huu + shift > 0.0) && shift < 1e30) shift *= 8.0;
            huu += shift;
        }
        const double ih = 1.0 / huu;
        El e;
        if (stage) {
            e.a00 = a00; e.a01 = a01; e.a10 = a10; e.a11 = a11;
            const double t = ih * gu;
            e.b0 = d0 - B0 * t; e.b1 = d1 - B1 * t;
            e.c00 = B0 * B0 * ih; e.c01 = B0 * B1 * ih; e.c11 = B1 * B1 * ih;
            e.q0 = w.GR[k * sG + 0]; e.q1 = w.GR[k * sG + 1];
            e.j00 = w.HS[k * sH + 0]; e.j01 = 0.0; e.j11 = w.HS[k * sH + 1];
        } else {
            e.a00 = e.a01 = e.a10 = e.a11 = e.b0 = e.b1 = e.c00 = e.c01 = e.c11 = 0.0;
            e.j00 = w.PP[N * nps + 0]; e.j01 = w.PP[N * nps + 1]; e.j11 = w.PP[N * nps + 2];
            e.q0 = w.PV[N * n + 0]; e.q1 = w.PV[N * n + 1];
        }
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            El r;
            shfl_el(tl, e, (lane + off) & (G - 1), r);
            if (lane + off <= N) {
                if (2 * off >= G) scan2::combine_jq(e, r, e.j00, e.j01, e.j11, e.q0, e.q1);
                else scan2::combine(e, r, e);
            }
        }
        const int nxt = (lane + 1) & (G - 1), prv = (lane - 1) & (G - 1);
        const double P00 = tl.shfl(e.j00, nxt), P01 = tl.shfl(e.j01, nxt), P11 = tl.shfl(e.j11, nxt);
        const double pv0 = tl.shfl(e.q0, nxt), pv1 = tl.shfl(e.q1, nxt);
        const double PB0 = P00 * B0 + P01 * B1, PB1 = P01 * B0 + P11 * B1;
        const double ihh = 1.0 / (huu + B0 * PB0 + B1 * PB1);
        const double Pd0 = P00 * d0 + P01 * d1 + pv0, Pd1 = P01 * d0 + P11 * d1 + pv1;
        const double K0 = -ihh * (PB0 * a00 + PB1 * a10), K1 = -ihh * (PB0 * a01 + PB1 * a11);
        const double kff = -ihh * (gu + B0 * Pd0 + B1 * Pd1);
        Af phi;                                              // closed-loop map of this stage (identity on the terminal lane)
        if (stage) {
            phi.m00 = a00 + B0 * K0; phi.m01 = a01 + B0 * K1; phi.m10 = a10 + B1 * K0; phi.m11 = a11 + B1 * K1;
            phi.v0 = B0 * kff + d0; phi.v1 = B1 * kff + d1;
        } else {
            phi.m00 = 1.0; phi.m01 = 0.0; phi.m10 = 0.0; phi.m11 = 1.0; phi.v0 = 0.0; phi.v1 = 0.0;
        }
        Af f = phi;
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            const int src = (lane - off) & (G - 1);
            Af r;
            r.m00 = tl.shfl(f.m00, src); r.m01 = tl.shfl(f.m01, src); r.m10 = tl.shfl(f.m10, src); r.m11 = tl.shfl(f.m11, src);
            r.v0 = tl.shfl(f.v0, src); r.v1 = tl.shfl(f.v1, src);
            if (lane >= off) scan2::compose(f, r, f);
        }
        double x0 = tl.shfl(f.v0, prv), x1 = tl.shfl(f.v1, prv);             // predictor dx_k
        if (lane == 0) { x0 = 0.0; x1 = 0.0; }
        const double du_a = K0 * x0 + K1 * x1 + kff;
        // ---- the predictor's slack and multiplier steps, its longest steps, the complementarity it would leave
        double lo, hi;
        M::bounds(prm, 0, lo, hi);
        const double sv = w.S[k * nr], zl = w.ZL[k * nr], zu = w.ZU[k * nr], isl = w.ISL[k * nr], isu = w.ISU[k * nr];
        const double sl = sv - lo, su = hi - sv;
        const double ds_a = M::row_sa(0) * du_a + w.RC[k * nr];
        const double dzl_a = -zl - zl * isl * ds_a, dzu_a = -zu + zu * isu * ds_a;
        double rp = 0.0, rd = 0.0, comp = 0.0;
        if (stage) {
            rp = dmax(-ds_a * isl, ds_a * isu);
            rd = dmax(1.0 + isl * ds_a, 1.0 - isu * ds_a);      // -dz_l / z_l, -dz_u / z_u of the predictor, without the divisions
            comp = zl * sl + zu * su;
        }
        rp = tl.max(dmax(rp, 0.0));
        rd = tl.max(dmax(rd, 0.0));
        comp = tl.sum(comp);
        const double apa = (rp > 1.0) ? 1.0 / rp : 1.0, ada = (rd > 1.0) ? 1.0 / rd : 1.0;
        double caff = 0.0;
        if (stage) caff = (sl + apa * ds_a) * (zl + ada * dzl_a) + (su - apa * ds_a) * (zu + ada * dzu_a);
        caff = tl.sum(caff);
        const double ratio = caff / comp;
        const double sigma = dmin(1.0, dmax(1e-8, ratio * ratio * ratio));
        const double mu = dmax(mu_min, sigma * comp * (0.5 / (double)N));      // mean over the 2 N bound pairs
        // ---- corrector: change of the input gradient against the predictor's, then the two affine recursions
        const double cl = ds_a * dzl_a, cu = ds_a * dzu_a;
        const double dg = stage ? M::row_sa(0) * ((mu + cu) * isu - (mu - cl) * isl) : 0.0;
        Af b;                                                // p -> Phi' p + K' dg, suffix composition
        b.m00 = phi.m00; b.m01 = phi.m10; b.m10 = phi.m01; b.m11 = phi.m11;
        b.v0 = stage ? K0 * dg : 0.0; b.v1 = stage ? K1 * dg : 0.0;
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            const int src = (lane + off) & (G - 1);
            Af r;
            r.m00 = tl.shfl(b.m00, src); r.m01 = tl.shfl(b.m01, src); r.m10 = tl.shfl(b.m10, src); r.m11 = tl.shfl(b.m11, src);
            r.v0 = tl.shfl(b.v0, src); r.v1 = tl.shfl(b.v1, src);
            if (lane + off <= N) scan2::compose(b, r, b);     // this lane's stages are applied last
        }
        const double dp0 = tl.shfl(b.v0, nxt), dp1 = tl.shfl(b.v1, nxt);      // dp_{k+1} (0 from the terminal lane)
        const double dkff = -ihh * (dg + B0 * dp0 + B1 * dp1);
        Af c;
        c.m00 = phi.m00; c.m01 = phi.m01; c.m10 = phi.m10; c.m11 = phi.m11;
        c.v0 = stage ? B0 * dkff : 0.0; c.v1 = stage ? B1 * dkff : 0.0;
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            const int src = (lane - off) & (G - 1);
            Af r;
            r.m00 = tl.shfl(c.m00, src); r.m01 = tl.shfl(c.m01, src); r.m10 = tl.shfl(c.m10, src); r.m11 = tl.shfl(c.m11, src);
            r.v0 = tl.shfl(c.v0, src); r.v1 = tl.shfl(c.v1, src);
            if (lane >= off) scan2::compose(c, r, c);
        }
        double y0 = tl.shfl(c.v0, prv), y1 = tl.shfl(c.v1, prv);             // correction of dx_k
        if (lane == 0) { y0 = 0.0; y1 = 0.0; w.DX[0] = 0.0; w.DX[1] = 0.0; }
        if (stage) {
            w.DU[k] = du_a + K0 * y0 + K1 * y1 + dkff;
            w.DX[(k + 1) * n + 0] = f.v0 + c.v0; w.DX[(k + 1) * n + 1] = f.v1 + c.v1;
            CL()[k] = cl; CU()[k] = cu;
            if (k + 1 < N) {                                 // P_N, p_N are already in place (prep)
                w.PP[(k + 1) * nps + 0] = P00; w.PP[(k + 1) * nps + 1] = P01; w.PP[(k + 1) * nps + 2] = P11;
                w.PV[(k + 1) * n + 0] = pv0 + dp0; w.PV[(k + 1) * n + 1] = pv1 + dp1;
            }
        }
        tl.sync();
        return mu;
    }

    // ---- the same sweeps with TWO consecutive elements per lane (lanes == (N + 1) / 2; four problems per warp): the lane
    // combines its two elements, the lane totals are scanned (log2(lanes) = 3 levels), and one reduced combination gives
    // the value function between the lane's two stages -- 5 combinations per lane for 4 problems per warp instead of
    // 4 for 2, i.e. 37 % fewer FP64 instructions per problem; the FP64 pipe's issue rate is what bounds this phase.
    struct StageLQ { double a00, a01, a10, a11, B0, B1, d0, d1, gu, huu, ih; };
    DART_HD void load_stage(int k, StageLQ& q) const {
        q.a00 = Aat(k, 0, 0); q.a01 = Aat(k, 0, 1); q.a10 = Aat(k, 1, 0); q.a11 = Aat(k, 1, 1);
        q.B0 = w.Bm[k * sB + 0]; q.B1 = w.Bm[k * sB + 1];
        q.d0 = w.D[k * sD + 0]; q.d1 = w.D[k * sD + 1];
        q.gu = w.GR[k * sG + 2];
        double huu = w.HS[k * sH + 2];
        if (!(huu > 0.0)) {
            double shift = 1e-4;
            while (!(huu + shift > 0.0) && shift < 1e30) shift *= 8.0;
            huu += shift;
        }
        q.huu = huu;
        q.ih = 1.0 / huu;
    }
    DART_HD void stage_element(int k, const StageLQ& q, scan2::El& e) const {
        e.a00 = q.a00; e.a01 = q.a01; e.a10 = q.a10; e.a11 = q.a11;
        const double t = q.ih * q.gu;
        e.b0 = q.d0 - q.B0 * t; e.b1 = q.d1 - q.B1 * t;
        e.c00 = q.B0 * q.B0 * q.ih; e.c01 = q.B0 * q.B1 * q.ih; e.c11 = q.B1 * q.B1 * q.ih;
        e.q0 = w.GR[k * sG + 0]; e.q1 = w.GR[k * sG + 1];
        e.j00 = w.HS[k * sH + 0]; e.j01 = 0.0; e.j11 = w.HS[k * sH + 1];
    }
    DART_HD void terminal_element(scan2::El& e) const {
        e.a00 = e.a01 = e.a10 = e.a11 = e.b0 = e.b1 = e.c00 = e.c01 = e.c11 = 0.0;
        e.j00 = w.PP[N * nps + 0]; e.j01 = w.PP[N * nps + 1]; e.j11 = w.PP[N * nps + 2];
        e.q0 = w.PV[N * n + 0]; e.q1 = w.PV[N * n + 1];
    }
    // gains of one stage from the next stage's value function; also the closed-loop map
    DART_HD static void stage_gains(const StageLQ& q, double P00, double P01, double P11, double pv0, double pv1, double& K0,
                                    double& K1, double& kff, scan2::Af& f) {
        double ihh;
        stage_gains(q, P00, P01, P11, pv0, pv1, K0, K1, kff, f, ihh);
    }
    DART_HD static void stage_gains(const StageLQ& q, double P00, double P01, double P11, double pv0, double pv1, double& K0,
                                    double& K1, double& kff, scan2::Af& f, double& ihh) {
        const double PB0 = P00 * q.B0 + P01 * q.B1, PB1 = P01 * q.B0 + P11 * q.B1;
        ihh = 1.0 / (q.huu + q.B0 * PB0 + q.B1 * PB1);
        const double Pd0 = P00 * q.d0 + P01 * q.d1 + pv0, Pd1 = P01 * q.d0 + P11 * q.d1 + pv1;
        K0 = -ihh * (PB0 * q.a00 + PB1 * q.a10); K1 = -ihh * (PB0 * q.a01 + PB1 * q.a11);
        kff = -ihh * (q.gu + q.B0 * Pd0 + q.B1 * Pd1);
        f.m00 = q.a00 + q.B0 * K0; f.m01 = q.a01 + q.B0 * K1; f.m10 = q.a10 + q.B1 * K0; f.m11 = q.a11 + q.B1 * K1;
        f.v0 = q.B0 * kff + q.d0; f.v1 = q.B1 * kff + q.d1;
    }
    template <class TL>
    DART_HD static void shfl_el(const TL& tl, const scan2::El& e, int src, scan2::El& r) {
        r.a00 = tl.shfl(e.a00, src); r.a01 = tl.shfl(e.a01, src); r.a10 = tl.shfl(e.a10, src); r.a11 = tl.shfl(e.a11, src);
        r.b0 = tl.shfl(e.b0, src); r.b1 = tl.shfl(e.b1, src);
        r.c00 = tl.shfl(e.c00, src); r.c01 = tl.shfl(e.c01, src); r.c11 = tl.shfl(e.c11, src);
        r.q0 = tl.shfl(e.q0, src); r.q1 = tl.shfl(e.q1, src);
        r.j00 = tl.shfl(e.j00, src); r.j01 = tl.shfl(e.j01, src); r.j11 = tl.shfl(e.j11, src);
    }
    template <class TL>
    DART_HD void sweeps_scan2(const TL& tl) {
        using scan2::El;
        using scan2::Af;
        constexpr int G = TL::kLanes;
        const int lane = tl.lane();
        const int kA = 2 * lane, kB = 2 * lane + 1;          // kA < N always; kB == N on the last lane (terminal element)
        const bool hasB = kB < N;
        StageLQ sa, sb;
        load_stage(kA, sa);
        load_stage(hasB ? kB : N - 1, sb);
        El eA, eB;
        stage_element(kA, sa, eA);
        if (hasB) stage_element(kB, sb, eB); else terminal_element(eB);
        El t;
        scan2::combine(eA, eB, t);
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            El r;
            shfl_el(tl, t, (lane + off) & (G - 1), r);
            if (lane + off < G) {
                if (2 * off >= G) scan2::combine_jq(t, r, t.j00, t.j01, t.j11, t.q0, t.q1);   // last level: only P, p are still needed
                else scan2::combine(t, r, t);
            }
        }
        // t = value function element of stage kA .. end; the next lane's t is that of stage kB + 1 .. end
        El nx;
        shfl_el(tl, t, (lane + 1) & (G - 1), nx);
        // value function after stage kA (= at kB): eB (+) nx, or the terminal element itself on the last lane
        double Pa00 = eB.j00, Pa01 = eB.j01, Pa11 = eB.j11, pa0 = eB.q0, pa1 = eB.q1;
        if (hasB) scan2::combine_jq(eB, nx, Pa00, Pa01, Pa11, pa0, pa1);
        double KA0, KA1, kfA, KB0 = 0.0, KB1 = 0.0, kfB = 0.0;
        Af fA, fB;
        stage_gains(sa, Pa00, Pa01, Pa11, pa0, pa1, KA0, KA1, kfA, fA);
        if (hasB) stage_gains(sb, nx.j00, nx.j01, nx.j11, nx.q0, nx.q1, KB0, KB1, kfB, fB);
        else { fB.m00 = 1.0; fB.m01 = 0.0; fB.m10 = 0.0; fB.m11 = 1.0; fB.v0 = 0.0; fB.v1 = 0.0; }
        Af F;
        scan2::compose(fB, fA, F);
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            const int src = (lane - off) & (G - 1);
            Af r;
            r.m00 = tl.shfl(F.m00, src); r.m01 = tl.shfl(F.m01, src); r.m10 = tl.shfl(F.m10, src); r.m11 = tl.shfl(F.m11, src);
            r.v0 = tl.shfl(F.v0, src); r.v1 = tl.shfl(F.v1, src);
            if (lane >= off) scan2::compose(F, r, F);
        }
        const int prv = (lane - 1) & (G - 1);
        double xa0 = tl.shfl(F.v0, prv), xa1 = tl.shfl(F.v1, prv);          // dx at stage kA
        if (lane == 0) { xa0 = 0.0; xa1 = 0.0; w.DX[0] = 0.0; w.DX[1] = 0.0; }
        const double xb0 = fA.m00 * xa0 + fA.m01 * xa1 + fA.v0, xb1 = fA.m10 * xa0 + fA.m11 * xa1 + fA.v1;   // dx at kB
        w.DU[kA] = KA0 * xa0 + KA1 * xa1 + kfA;
        w.DX[kB * n + 0] = xb0; w.DX[kB * n + 1] = xb1;
        if (hasB) {
            w.PP[kB * nps + 0] = Pa00; w.PP[kB * nps + 1] = Pa01; w.PP[kB * nps + 2] = Pa11;
            w.PV[kB * n + 0] = pa0; w.PV[kB * n + 1] = pa1;
            w.DU[kB] = KB0 * xb0 + KB1 * xb1 + kfB;
            w.DX[(kB + 1) * n + 0] = F.v0; w.DX[(kB + 1) * n + 1] = F.v1;
            if (kB + 1 < N) {
                w.PP[(kB + 1) * nps + 0] = nx.j00; w.PP[(kB + 1) * nps + 1] = nx.j01; w.PP[(kB + 1) * nps + 2] = nx.j11;
                w.PV[(kB + 1) * n + 0] = nx.q0; w.PV[(kB + 1) * n + 1] = nx.q1;
            }
        }
        tl.sync();
    }

    // ---- the predictor-corrector step (see sweeps_scan_pc) with two consecutive stages per lane
    template <class TL>
    DART_HD static void shfl_af(const TL& tl, const scan2::Af& f, int src, scan2::Af& r) {
        r.m00 = tl.shfl(f.m00, src); r.m01 = tl.shfl(f.m01, src); r.m10 = tl.shfl(f.m10, src); r.m11 = tl.shfl(f.m11, src);
        r.v0 = tl.shfl(f.v0, src); r.v1 = tl.shfl(f.v1, src);
    }
    struct RowPC { double sl, su, zl, zu, isl, isu, ds, dzl, dzu; };
    DART_HD void row_affine(int k, double du, double lo, double hi, RowPC& r) const {
        const double sv = w.S[k * nr];
        r.zl = w.ZL[k * nr]; r.zu = w.ZU[k * nr]; r.isl = w.ISL[k * nr]; r.isu = w.ISU[k * nr];
        r.sl = sv - lo; r.su = hi - sv;
        r.ds = M::row_sa(0) * du + w.RC[k * nr];
        r.dzl = -r.zl - r.zl * r.isl * r.ds; r.dzu = -r.zu + r.zu * r.isu * r.ds;
    }
    template <class TL>
    DART_HD double sweeps_scan2_pc(const TL& tl, double mu_min) {
        static_assert(nr == 1 && m == 1 && n == 2, "predictor-corrector step: 2-state / 1-input problems with one bound row");
        using scan2::El;
        using scan2::Af;
        constexpr int G = TL::kLanes;
        const int lane = tl.lane();
        const int kA = 2 * lane, kB = 2 * lane + 1;
        const bool hasB = kB < N;
        const int kBs = hasB ? kB : N - 1;                   // stage the B slot reads when it is the terminal element (unused values)
        StageLQ sa, sb;
        load_stage(kA, sa);
        load_stage(kBs, sb);
        El eA, eB;
        stage_element(kA, sa, eA);
        if (hasB) stage_element(kB, sb, eB); else terminal_element(eB);
        El t;
        scan2::combine(eA, eB, t);
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            El r;
            shfl_el(tl, t, (lane + off) & (G - 1), r);
            if (lane + off < G) {
                if (2 * off >= G) scan2::combine_jq(t, r, t.j00, t.j01, t.j11, t.q0, t.q1);
                else scan2::combine(t, r, t);
            }
        }
        const int nxt = (lane + 1) & (G - 1), prv = (lane - 1) & (G - 1);
        const bool last = lane == G - 1;
        El nx;
        shfl_el(tl, t, nxt, nx);
        double Pa00 = eB.j00, Pa01 = eB.j01, Pa11 = eB.j11, pa0 = eB.q0, pa1 = eB.q1;
        if (hasB) scan2::combine_jq(eB, nx, Pa00, Pa01, Pa11, pa0, pa1);
        double KA0, KA1, kfA, ihA, KB0 = 0.0, KB1 = 0.0, kfB = 0.0, ihB = 0.0;
        Af fA, fB;
        stage_gains(sa, Pa00, Pa01, Pa11, pa0, pa1, KA0, KA1, kfA, fA, ihA);
        if (hasB) stage_gains(sb, nx.j00, nx.j01, nx.j11, nx.q0, nx.q1, KB0, KB1, kfB, fB, ihB);
        else { fB.m00 = 1.0; fB.m01 = 0.0; fB.m10 = 0.0; fB.m11 = 1.0; fB.v0 = 0.0; fB.v1 = 0.0; }
        Af F;
        scan2::compose(fB, fA, F);
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            Af r;
            shfl_af(tl, F, (lane - off) & (G - 1), r);
            if (lane >= off) scan2::compose(F, r, F);
        }
        double xa0 = tl.shfl(F.v0, prv), xa1 = tl.shfl(F.v1, prv);          // predictor dx at stage kA
        if (lane == 0) { xa0 = 0.0; xa1 = 0.0; }
        const double xb0 = fA.m00 * xa0 + fA.m01 * xa1 + fA.v0, xb1 = fA.m10 * xa0 + fA.m11 * xa1 + fA.v1;
        const double duA = KA0 * xa0 + KA1 * xa1 + kfA, duB = KB0 * xb0 + KB1 * xb1 + kfB;
        // ---- predictor's slack / multiplier steps, longest steps, complementarity left
        double lo, hi;
        M::bounds(prm, 0, lo, hi);
        RowPC ra, rb;
        row_affine(kA, duA, lo, hi, ra);
        row_affine(kBs, duB, lo, hi, rb);
        double rp = dmax(-ra.ds * ra.isl, ra.ds * ra.isu), rd = dmax(1.0 + ra.isl * ra.ds, 1.0 - ra.isu * ra.ds);   // -dz / z of the predictor, division-free
        double comp = ra.zl * ra.sl + ra.zu * ra.su;
        if (hasB) {
            rp = dmax(rp, dmax(-rb.ds * rb.isl, rb.ds * rb.isu));
            rd = dmax(rd, dmax(1.0 + rb.isl * rb.ds, 1.0 - rb.isu * rb.ds));
            comp += rb.zl * rb.sl + rb.zu * rb.su;
        }
        rp = tl.max(dmax(rp, 0.0));
        rd = tl.max(dmax(rd, 0.0));
        comp = tl.sum(comp);
        const double apa = (rp > 1.0) ? 1.0 / rp : 1.0, ada = (rd > 1.0) ? 1.0 / rd : 1.0;
        double caff = (ra.sl + apa * ra.ds) * (ra.zl + ada * ra.dzl) + (ra.su - apa * ra.ds) * (ra.zu + ada * ra.dzu);
        if (hasB) caff += (rb.sl + apa * rb.ds) * (rb.zl + ada * rb.dzl) + (rb.su - apa * rb.ds) * (rb.zu + ada * rb.dzu);
        caff = tl.sum(caff);
        const double ratio = caff / comp;
        const double sigma = dmin(1.0, dmax(1e-8, ratio * ratio * ratio));
        const double mu = dmax(mu_min, sigma * comp * (0.5 / (double)N));
        // ---- corrector
        const double clA = ra.ds * ra.dzl, cuA = ra.ds * ra.dzu, clB = rb.ds * rb.dzl, cuB = rb.ds * rb.dzu;
        const double dgA = M::row_sa(0) * ((mu + cuA) * ra.isu - (mu - clA) * ra.isl);
        const double dgB = hasB ? M::row_sa(0) * ((mu + cuB) * rb.isu - (mu - clB) * rb.isl) : 0.0;
        // backward: p -> Phi' p + K' dg per stage; the lane's map applies stage kB first, then kA
        Af bA, bB, bt;
        bA.m00 = fA.m00; bA.m01 = fA.m10; bA.m10 = fA.m01; bA.m11 = fA.m11; bA.v0 = KA0 * dgA; bA.v1 = KA1 * dgA;
        bB.m00 = fB.m00; bB.m01 = fB.m10; bB.m10 = fB.m01; bB.m11 = fB.m11; bB.v0 = KB0 * dgB; bB.v1 = KB1 * dgB;
        scan2::compose(bA, bB, bt);
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            Af r;
            shfl_af(tl, bt, (lane + off) & (G - 1), r);
            if (lane + off < G) scan2::compose(bt, r, bt);
        }
        double q0 = tl.shfl(bt.v0, nxt), q1 = tl.shfl(bt.v1, nxt);          // dp at stage kB + 1 (0 past the horizon)
        if (last) { q0 = 0.0; q1 = 0.0; }
        const double dpB0 = bB.m00 * q0 + bB.m01 * q1 + bB.v0, dpB1 = bB.m10 * q0 + bB.m11 * q1 + bB.v1;   // dp at stage kB
        const double dkB = hasB ? -ihB * (dgB + sb.B0 * q0 + sb.B1 * q1) : 0.0;
        const double dkA = -ihA * (dgA + sa.B0 * dpB0 + sa.B1 * dpB1);
        // forward: dx -> Phi dx + B dkff per stage
        Af cA, cB, ct;
        cA.m00 = fA.m00; cA.m01 = fA.m01; cA.m10 = fA.m10; cA.m11 = fA.m11; cA.v0 = sa.B0 * dkA; cA.v1 = sa.B1 * dkA;
        cB.m00 = fB.m00; cB.m01 = fB.m01; cB.m10 = fB.m10; cB.m11 = fB.m11; cB.v0 = hasB ? sb.B0 * dkB : 0.0; cB.v1 = hasB ? sb.B1 * dkB : 0.0;
        scan2::compose(cB, cA, ct);
        DART_UNROLL for (int off = 1; off < G; off <<= 1) {
            Af r;
            shfl_af(tl, ct, (lane - off) & (G - 1), r);
            if (lane >= off) scan2::compose(ct, r, ct);
        }
        double ya0 = tl.shfl(ct.v0, prv), ya1 = tl.shfl(ct.v1, prv);        // correction of dx at stage kA
        if (lane == 0) { ya0 = 0.0; ya1 = 0.0; w.DX[0] = 0.0; w.DX[1] = 0.0; }
        const double yb0 = cA.m00 * ya0 + cA.m01 * ya1 + cA.v0, yb1 = cA.m10 * ya0 + cA.m11 * ya1 + cA.v1;
        w.DU[kA] = duA + KA0 * ya0 + KA1 * ya1 + dkA;
        w.DX[kB * n + 0] = xb0 + yb0; w.DX[kB * n + 1] = xb1 + yb1;
        CL()[kA] = clA; CU()[kA] = cuA;
        if (hasB) {
            w.PP[kB * nps + 0] = Pa00; w.PP[kB * nps + 1] = Pa01; w.PP[kB * nps + 2] = Pa11;
            w.PV[kB * n + 0] = pa0 + dpB0; w.PV[kB * n + 1] = pa1 + dpB1;
            w.DU[kB] = duB + KB0 * yb0 + KB1 * yb1 + dkB;
            w.DX[(kB + 1) * n + 0] = F.v0 + ct.v0; w.DX[(kB + 1) * n + 1] = F.v1 + ct.v1;
            CL()[kB] = clB; CU()[kB] = cuB;
            if (kB + 1 < N) {
                w.PP[(kB + 1) * nps + 0] = nx.j00; w.PP[(kB + 1) * nps + 1] = nx.j01; w.PP[(kB + 1) * nps + 2] = nx.j11;
                w.PV[(kB + 1) * n + 0] = nx.q0 + q0; w.PV[(kB + 1) * n + 1] = nx.q1 + q1;
            }
        }
        tl.sync();
        return mu;
    }

    // ---- slack steps, step-length limits, directional derivative, then the dual step z += ad dz (stage-parallel).
    // dz is recomputed in the second pass instead of being stored; the new equality multipliers are formed in move_dual.
    // pc: the direction is the corrected one of sweeps_scan_pc -- the multiplier steps aim at mu -/+ the predictor's products
    DART_HD void post(double mu, double& ap, double& ad, double& dphi, bool ghost = false, bool pc = false) {
        const double tau = dmax(o.tau_min, 1.0 - mu);
        double ap_ = 1.0, ad_ = 1.0, dp = 0.0, rp_ = 0.0, rd_ = 0.0;
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int i = 0; i < ny; ++i) {
                double d = (i < n) ? w.DX[k * n + i] : w.DU[k * m + (i - n)];
                dp += cost_grad(k, i) * d;
            }
            double zz[nr], na[nr], nb[nr];                 // z_l z_u and the numerators of -dz_l/z_l, -dz_u/z_u
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) { w.DS[k * nr + r] = 0.0; zz[r] = 1.0; na[r] = 0.0; nb[r] = 0.0; continue; }
                const int ia = M::row_ia(r), ib = M::row_ib(r);
                double dy = M::row_sa(r) * ((ia < n) ? w.DX[k * n + ia] : w.DU[k * m + (ia - n)]);
                if (ib >= 0) dy += M::row_sb(r) * ((ib < n) ? w.DX[k * n + ib] : w.DU[k * m + (ib - n)]);
                const double ds = dy + w.RC[k * nr + r];
                const double isl = w.ISL[k * nr + r], isu = w.ISU[k * nr + r];
                const double zl = w.ZL[k * nr + r], zu = w.ZU[k * nr + r];
                double ml = mu, mu_u = mu;
                if (pc) {
                    if constexpr (kVec) {
                        // tiled path (pc_rows): DS holds the predictor's row step; it moves to RC, which is dead until
                        // the line search's eval1 rewrites it, for the second pass below
                        const double dsa = w.DS[k * nr + r];
                        w.RC[k * nr + r] = dsa;
                        ml = mu - dsa * (-zl - zl * isl * dsa);
                        mu_u = mu + dsa * (-zu + zu * isu * dsa);
                    } else {
                        ml = mu - CL()[k * nr + r];
                        mu_u = mu + CU()[k * nr + r];
                    }
                }
                const double dzl = ml * isl - zl - zl * isl * ds;
                const double dzu = mu_u * isu - zu + zu * isu * ds;
                w.DS[k * nr + r] = ds;
                dp -= mu * ds * (isl - isu);
                rp_ = dmax(rp_, dmax(-ds * isl, ds * isu));
                zz[r] = zl * zu; na[r] = -dzl * zu; nb[r] = -dzu * zl;
            }
            double iz[nr];
            batch_inv(zz, iz);                             // one division per stage
            DART_UNROLL for (int r = 0; r < nr; ++r) rd_ = dmax(rd_, dmax(na[r] * iz[r], nb[r] * iz[r]));
        }
        if (tile.lane() == 0) {
            DART_UNROLL for (int i = 0; i < n; ++i)
                dp += 2.0 * M::wT(prm, i) * (w.X[N * n + i] - M::rT(prm, w.REF, N, i)) * w.DX[N * n + i];
        }
        rp_ = tile.max_nonneg(rp_);
        rd_ = tile.max_nonneg(rd_);
        ap_ = (rp_ > tau) ? tau / rp_ : 1.0;
        ad_ = (rd_ > tau) ? tau / rd_ : 1.0;
        if (ghost) { ap_ = 0.0; ad_ = 0.0; }      // a finished tile in lockstep with its sibling: nothing moves
        ap = ap_;
        ad = ad_;
        dphi = tile.sum(dp);
        if (ad_ != 0.0)
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) continue;
                const double ds = w.DS[k * nr + r];
                const double isl = w.ISL[k * nr + r], isu = w.ISU[k * nr + r];
                const double zl = w.ZL[k * nr + r], zu = w.ZU[k * nr + r];
                double ml = mu, mu_u = mu;
                if (pc) {
                    if constexpr (kVec) {
                        const double dsa = w.RC[k * nr + r];
                        ml = mu - dsa * (-zl - zl * isl * dsa);
                        mu_u = mu + dsa * (-zu + zu * isu * dsa);
                    } else {
                        ml = mu - CL()[k * nr + r];
                        mu_u = mu + CU()[k * nr + r];
                    }
                }
                w.ZL[k * nr + r] = zl + ad_ * (ml * isl - zl - zl * isl * ds);
                w.ZU[k * nr + r] = zu + ad_ * (mu_u * isu - zu + zu * isu * ds);
            }
        }
        tile.sync();
    }

    DART_HD void move_primal(double da) {
        if (da != 0.0)                      // a zero step (a tile kept in lockstep by the line search of its warp) moves nothing
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int i = 0; i < n; ++i) w.X[(k + 1) * n + i] += da * w.DX[(k + 1) * n + i];
            DART_UNROLL for (int j = 0; j < m; ++j) w.U[k * m + j] += da * w.DU[k * m + j];
            DART_UNROLL for (int r = 0; r < nr; ++r) w.S[k * nr + r] += da * w.DS[k * nr + r];
        }
        tile.sync();
    }

    DART_HD void move_dual(double alpha, double mu) {
        const double ks = 1e10, iks = 1e-10;     // safeguard interval [mu/(ks s), ks mu/s]; a product, not a division
        if (alpha != 0.0)                        // zero step: a finished / stopping tile kept in lockstep -- its multipliers stay
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int a = 0; a < n; ++a) {      // new multiplier of stage k+1: P_{k+1} dx_{k+1} + p_{k+1}
                double ln = w.PV[(k + 1) * n + a];
                DART_UNROLL for (int b = 0; b < n; ++b) ln += Pat(k + 1, a, b) * w.DX[(k + 1) * n + b];
                w.LAM[k * n + a] += alpha * (ln - w.LAM[k * n + a]);
            }
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) continue;
                const double isl = w.ISL[k * nr + r], isu = w.ISU[k * nr + r];   // of the accepted point (eval1)
                const double zl = w.ZL[k * nr + r], zu = w.ZU[k * nr + r];        // already stepped in post()
                w.ZL[k * nr + r] = dmin(dmax(zl, mu * isl * iks), ks * mu * isl);
                w.ZU[k * nr + r] = dmin(dmax(zu, mu * isu * iks), ks * mu * isu);
            }
        }
        tile.sync();
    }

    // ---- slack/dual initialisation (IPOPT bound_push / bound_frac; z = mu0 / slack)
    DART_HD void init_rows(double mu) {
        if (tile.lane() == 0) w.ZR[0] = 0.0;
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                double lo, hi;
                M::bounds(prm, r, lo, hi);
                double push = dmin(o.bound_push * dmax(1.0, dmax(fabs(lo), fabs(hi))), o.bound_push * (hi - lo));
                double s = dmin(dmax(rowval(k, r), lo + push), hi - push);
                if (M::row_ib(r) < 0 && M::row_ia(r) >= n) w.U[k * m + (M::row_ia(r) - n)] = s / M::row_sa(r);
                w.S[k * nr + r] = s;
                bool mk = masked(k, r);
                w.ZL[k * nr + r] = mk ? 0.0 : mu / (s - lo);
                w.ZU[k * nr + r] = mk ? 0.0 : mu / (hi - s);
            }
            DART_UNROLL for (int a = 0; a < n; ++a) w.LAM[k * n + a] = 0.0;
        }
        tile.sync();
        if (M::NAUG > 0) {
            for (int k = tile.lane(); k < N; k += tile.size())
                DART_UNROLL for (int j = 0; j < m; ++j) w.X[(k + 1) * n + np + j] = w.U[k * m + j];
            tile.sync();
        }
    }

    // ---- dual warm start: per-problem block [valid | LAM (N n) | S (N nr) | ZL | ZU] in global memory.
    // The previous optimum's slacks sit within ~1e-9 of their active bounds; they are pushed 1e-6 of the row's range
    // into the interior, and the bound multipliers are kept from falling below IPOPT's safeguard for the new mu.
    DART_HD static int dual_doubles(int N_) { return 1 + N_ * n + 3 * N_ * nr; }
    // `enabled` = false: take part in the tile syncs only (lockstep tiles: the sibling tile of the warp may be loading)
    DART_HD void load_duals(const double* blk, double mu, bool enabled = true) {
        const double* lam = blk + 1;
        const double* sp = lam + N * n;
        const double* zlp = sp + N * nr;
        const double* zup = zlp + N * nr;
        const double ks = 1e10;
        if (enabled)
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                if (masked(k, r)) continue;
                double lo, hi;
                M::bounds(prm, r, lo, hi);
                const double push = 1e-6 * (hi - lo);
                const double sv = dmin(dmax(sp[k * nr + r], lo + push), hi - push);
                if (M::row_ib(r) < 0 && M::row_ia(r) >= n) w.U[k * m + (M::row_ia(r) - n)] = sv / M::row_sa(r);
                w.S[k * nr + r] = sv;
                w.ZL[k * nr + r] = dmax(zlp[k * nr + r], mu / (ks * (sv - lo)));
                w.ZU[k * nr + r] = dmax(zup[k * nr + r], mu / (ks * (hi - sv)));
            }
            DART_UNROLL for (int a = 0; a < n; ++a) w.LAM[k * n + a] = lam[k * n + a];
        }
        tile.sync();
        if (M::NAUG > 0) {
            if (enabled)
            for (int k = tile.lane(); k < N; k += tile.size())
                DART_UNROLL for (int j = 0; j < m; ++j) w.X[(k + 1) * n + np + j] = w.U[k * m + j];
            tile.sync();
        }
    }
    DART_HD void store_duals(double* blk, bool usable) {
        double* lam = blk + 1;
        double* sp = lam + N * n;
        double* zlp = sp + N * nr;
        double* zup = zlp + N * nr;
        for (int k = tile.lane(); k < N; k += tile.size()) {
            DART_UNROLL for (int r = 0; r < nr; ++r) {
                sp[k * nr + r] = w.S[k * nr + r]; zlp[k * nr + r] = w.ZL[k * nr + r]; zup[k * nr + r] = w.ZU[k * nr + r];
            }
            DART_UNROLL for (int a = 0; a < n; ++a) lam[k * n + a] = w.LAM[k * n + a];
        }
        if (tile.lane() == 0) blk[0] = usable ? 1.0 : 0.0;
    }

    // ---- the interior-point loop.  X (all N+1 states, X[0] = x0), U and REF must be set by the caller when `active`.
    //
    // Every iteration has three phases with a phase-dependent mapping of work to threads:
    //   A (tile = G lanes of this problem): convergence test, barrier update, prep(); tiled Riccati sweep for the
    //     larger models;
    //   B (one THREAD per problem of the block): the serial Riccati/forward sweeps -- thread q of the block runs the
    //     sweeps of problem q out of its shared-memory workspace, so a warp advances 32 problems per instruction
    //     instead of replicating one recursion in every lane;
    //   C (tile): slack/dual steps, line search with re-evaluation, multiplier update, KKT residuals.
    // Two block barriers per iteration separate A|B|C; all tiles of the block (also finished or empty ones) keep
    // taking part in them until no problem of the block needs another sweep.
    // scan sweeps: 2-state / 1-input problems on a tile with exactly one lane per stage plus the terminal element
    // (the tile type says whether the tiles of a warp run in lockstep: nmpc_kernel.cuh TileFor)
    static constexpr bool kScan = M::SERIAL_RICCATI && n == 2 && m == 1 && NC > 0 && T::kLockstep &&
                                  (T::kLanes == NC + 1 || 2 * T::kLanes == NC + 1);
    // sub-warp tiles whose warp runs ONE instruction stream (collectives with a constant full-warp mask)
    static constexpr bool kLock = T::kLockstep && T::kLanes < 32;
    // Mehrotra predictor-corrector instead of the monotone barrier schedule: where the step is built by sweeps_scan_pc
    // (kPCT: the tiled-sweep models -- ordinary sweeps on the affine-scaling right-hand side, then pc_rows + corrector_*)
    static constexpr bool kPCT = kVec && M::MEHROTRA;
    static constexpr bool kPC = (kScan || kPCT) && M::MEHROTRA;
    DART_HD static bool uses_pc(const SolverOpts& oo) {
        // 2: DART_BARRIER_AUTO -- the method's default, or (PC_COLD) predictor-corrector steps for launches without a warm plan
        return kPC && (oo.mehrotra == 1 || (oo.mehrotra == 2 && (M::PC_DEFAULT || (M::PC_COLD && oo.cold != 0))));
    }
    // multiplier scale of a cold start, z = mu0 / slack: IPOPT's mu_init (0.1) under the monotone schedule, which starts its
    // barrier parameter there; predictor-corrector steps pick mu themselves, and a start closer to the final complementarity
    // saves them iterations (oracle, 96-288 instances: PMPC 5.79 -> 5.20, RMPC 9.41 -> 8.69, LMPC 7.10 -> 5.75 at 0.01): Model::MU0_PC,
    // used when mu_init was left at 0
    DART_HD static double start_mu(const SolverOpts& oo) { return (uses_pc(oo) && oo.mu0_auto) ? M::MU0_PC : oo.mu0; }

    // The caller has initialised slacks and multipliers (init_rows, optionally load_duals) for barrier parameter mu0.
    DART_HD void run(bool active, double mu0, double& J, int32_t& status, int32_t& iters, double& kkt) {
        double mu = mu0;
        double f = 0.0, L = 0.0, th = 0.0, pinf = 0.0, dinf = 0.0, zs_min = 0.0, zs_max = 0.0, lam_sum = 0.0, z_sum = 0.0;
        // number of constraint rows of the problem (rows with row_skip0 do not exist at stage 0)
        int nact = 0;
        DART_UNROLL for (int r = 0; r < nr; ++r) nact += M::row_skip0(r) ? N - 1 : N;
        double inv_nd = 0.0, inv_nc = 0.0;
        if (active) {
            inv_nd = 1.0 / (double)(N * n + 2 * nact);
            inv_nc = 1.0 / (double)(2 * (nact > 0 ? nact : 1));
        }
        // The starting point is evaluated by the FIRST PASS of the loop below, through the same two call sites of eval1 /
        // eval2 that the iterations use: every call site is an inlined copy of the RK4 + sensitivity code, the kernels are
        // 150-190 kB of SASS against a 32 kB instruction cache, and instruction fetch was the LMPC kernel's largest stall
        // reason (no_instruction 30 % of the stall samples).
        bool first = true;
        const double mu_min = o.tol / 10.0;
        int it = 0, tiny = 0, nacc = 0;
        int32_t st = ST_MAXITER;
        double E0 = 0.0, is_d = 1.0, is_c = 1.0;
        bool done = !active;
        double* myslot = bc.base + (size_t)(bc.tid / tile.size()) * bc.stride;
        const bool pc = uses_pc(o);
        // forward sweep of the tiled-sweep models; with predictor-corrector steps it runs twice through ONE copy of the code
        // (instruction cache): predictor, then -- after pc_rows and the corrector's backward vector recursion have replaced
        // the feed-forward terms -- the corrected step
        auto forward_sweeps = [&](bool ghost_) {
            DART_UNROLL_N(1)
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {
                    if constexpr (kPCT) {
                        if (!pc) break;
                        const double mu_new = pc_rows(mu_min, inv_nc);
                        if (!ghost_) mu = mu_new;
                        if constexpr (T::kLanes >= n) corrector_tile(tile);
                        else {
                            if (tile.lane() == 0) corrector_serial_backward();
                            tile.sync();
                        }
                        feedforward();
                    } else {
                        break;
                    }
                }
                if (T::kLanes >= n) forward_tile(tile);
                else if (tile.lane() == 0) forward();
                tile.sync();
            }
        };
        static_assert(!kLock || kScan || !M::SERIAL_RICCATI, "lockstep tiles need the scan sweeps or the tiled Riccati sweep");
#ifdef DART_PHASE_CLOCK
        long long ckA = 0, ckB = 0, ckC = 0, ckBb = 0, ckC1 = 0, ckC2 = 0, ckA1 = 0, ckW1 = 0, ckW2 = 0, ck0 = DART_CLOCK();
#define DART_CK(acc) { long long t_ = DART_CLOCK(); acc += t_ - ck0; ck0 = t_; }
#else
#define DART_CK(acc)
#endif
        for (;;) {
            bool ghost = false;
            double ap = 0.0, ad = 0.0, dphi = 0.0;
            if (!first) {
            // ---------------- phase A
            bool need_sweep = false;
            if (!done) {
                // IPOPT's scaled optimality error; is_d = 1/s_d, is_c = 1/s_c (s_* = max(smax, mean multiplier)/smax)
                // (the scalings are exactly 1 unless the mean multiplier exceeds smax: no division on the usual path)
                const double md = (lam_sum + z_sum) * inv_nd, mc = z_sum * inv_nc;
                is_d = (md > o.smax) ? o.smax / md : 1.0;
                is_c = (mc > o.smax) ? o.smax / mc : 1.0;
                const double base = dmax(dinf * is_d, pinf);
                E0 = dmax(base, (nact > 0 ? zs_max : 0.0) * is_c);
                if (E0 <= o.tol) { st = ST_CONVERGED; done = true; }
                else if (!(E0 == E0) || E0 > 1e300) { st = ST_NUMERIC; done = true; }
                else {
                    nacc = (o.acc_iter > 0 && E0 <= o.acc_tol) ? nacc + 1 : 0;
                    if (o.acc_iter > 0 && nacc >= o.acc_iter) { st = ST_ACCEPTABLE; done = true; }
                    else if (it >= o.max_iter) done = true;
                }
                if (!done) {
                    if (!pc)
                    for (int q = 0; q < 8; ++q) {
                        double cm = (nact > 0) ? dmax(fabs(zs_max - mu), fabs(zs_min - mu)) : 0.0;
                        double Emu = dmax(base, cm * is_c);
                        if (Emu <= o.kappa_eps * mu && mu > mu_min)
                            mu = dmax(mu_min, dmin(o.kappa_mu * mu, (o.theta_mu == 1.5) ? mu * sqrt(mu) : pow(mu, o.theta_mu)));
                        else break;
                    }
                    DART_CK(ckA1)
                    need_sweep = true;
                    if constexpr (!kLock) {
                        prep((kPCT && pc) ? 0.0 : mu);
                        if (!M::SERIAL_RICCATI) backward(tile);
                    }
                }
            }
            if constexpr (kLock) {
                // Lockstep tiles: the tiles of a WARP run the whole iteration as one instruction stream, so that every
                // collective is a full-warp instruction with a constant mask.  A tile that has finished while a sibling
                // has not keeps executing the same instructions as a "ghost" -- zero step lengths, so its primal point,
                // objective and status stay exactly what they were when it finished.  No block barriers: warps are
                // independent.  (PMPC axis problems sweep by scans, the larger models by the tiled Riccati sweep.)
                DART_CK(ckA)
                if (tile.warp_ballot(need_sweep) == 0u) break;
                ghost = !need_sweep;
                bool stepped = false;
                if constexpr (kPC && kScan) {
                    if (pc) {
                        prep(0.0);                                   // affine-scaling right-hand side
                        double mu_new;
                        if constexpr (T::kLanes == NC + 1) mu_new = sweeps_scan_pc(tile, mu_min); else mu_new = sweeps_scan2_pc(tile, mu_min);
                        if (!ghost) mu = mu_new;
                        stepped = true;
                    }
                }
                if (!stepped) prep((kPCT && pc) ? 0.0 : mu);
                DART_CK(ckW1)
                if constexpr (kScan) {
                    if (!stepped) { if constexpr (T::kLanes == NC + 1) sweeps_scan(tile); else sweeps_scan2(tile); }
                } else {
                    backward(tile);
                    forward_sweeps(ghost);
                }
                DART_CK(ckB)
            } else if (M::SERIAL_RICCATI) {
                if (tile.lane() == 0) myslot[0] = need_sweep ? 1.0 : 0.0;
                DART_CK(ckA)
                const bool any_ = tile.block_any(need_sweep);   // barrier + vote: phase-A writes are visible past this point
                DART_CK(ckW1)
                if (!any_) break;
                // ---------------- phase B: thread q sweeps problem q
                if (bc.tid < bc.nprob) {
                    double* sl = bc.base + (size_t)bc.tid * bc.stride;
                    if (sl[0] != 0.0) {
                        W wk;
                        wk.bind(sl + kSlot, N);
                        Solver other(tile, prm, o, N, wk, bc);
                        other.backward_serial();
                        DART_CK(ckBb)
                        other.forward();
                    }
                }
                DART_CK(ckB)
                tile.block_sync();
                DART_CK(ckW2)
                if (!need_sweep) continue;
            } else {
                // larger models: the tile swept backwards itself; its problems are independent of the rest of the block,
                // so no block barrier -- one lane runs the short forward recurrence
                DART_CK(ckA)
                if (!need_sweep) break;
                forward_sweeps(false);
                DART_CK(ckB)
            }
            // ---------------- phase C
            post(mu, ap, ad, dphi, ghost, pc);
            DART_CK(ckC1)
            }   // !first
            const double phi0 = f - mu * L, th0 = th;
            const double th_max = 1e4 * dmax(1.0, th0);
            double alpha = ap, applied = 0.0;
            for (int bt = 0;; ++bt) {
                if (!first) {
                    move_primal(alpha - applied);
                    applied = alpha;
                }
                if (!first || active) eval1(f, L, th, pinf);
                if (first) break;
                const double phit = f - mu * L;
                // alpha |dphi|^s_phi > delta th0^s_theta, compared in the log domain (only needed when th0 is small);
                // single precision is ample for this heuristic test (and keeps three FP64 logs off the serial path)
                bool switching = false;
                if (dphi < 0.0 && th0 <= o.theta_small)
                    switching = (th0 <= 0.0) || (log2f((float)alpha) + (float)o.s_phi * log2f((float)(-dphi)) >
                                                 log2f((float)o.delta_sw) + (float)o.s_theta * log2f((float)th0));
                bool armijo = phit <= phi0 + o.eta * alpha * dphi + 10.0 * 2.220446049250313e-16 * fabs(phi0);
                bool suff = (th <= (1.0 - o.gamma_theta) * th0) || (phit <= phi0 - o.gamma_phi * th0);
                bool ok = ((switching && th0 <= o.theta_small) ? armijo : suff) && (th <= th_max) && (phit == phit) &&
                          (fabs(phit) < 1e300);
                const bool stop_ls = ok || bt + 1 >= o.max_backtrack || ghost;
                if constexpr (kLock) {
                    // the tiles of a warp leave the line search together (a tile that is done re-evaluates its accepted
                    // point with a zero step: same values), so the collectives inside stay warp-converged
                    if (tile.group_all(stop_ls)) break;
                    if (!stop_ls) alpha *= 0.5;
                } else {
                    if (stop_ls) break;
                    alpha *= 0.5;
                }
            }
#ifdef DART_TRACE
            printf("it %d mu %.3e E0 %.3e dinf %.3e pinf %.3e zsmax %.3e ap %.4f ad %.4f alpha %.5f dphi %.3e th0 %.3e f %.10g\n", it, mu, E0, dinf * is_d, pinf, zs_max * is_c, ap, ad, alpha, dphi, th0, f);
#endif
            DART_CK(ckC2)
            if (!first) {
                if (!ghost) ++it;
                // the step vanished three times in a row: no restoration phase here -- stop and say so
                tiny = ghost ? 0 : ((alpha <= 1e-6) ? tiny + 1 : 0);
                if (tiny >= 3) {
                    st = (pinf > 1e-4) ? ST_INFEASIBLE : ST_MAXITER;
                    done = true;
                    if constexpr (!kLock) continue;
                    alpha = 0.0;                      // lockstep tiles: stay with the other tiles of the warp (their collectives below)
                }
                move_dual(alpha, mu);
            }
            if (!first || active) eval2(dinf, zs_min, zs_max, lam_sum, z_sum);
            first = false;
            DART_CK(ckC)
        }
#ifdef DART_PHASE_CLOCK
#ifdef __CUDA_ARCH__
        if (bc.tid == 0 && DART_PHASE_CLOCK_BLOCK)
            printf("sub-phases per iter: A.test %lld  B.backward %lld  C.post %lld  C.linesearch %lld\n", ckA1 / (it + 1), ckBb / (it > 0 ? it : 1), ckC1 / (it > 0 ? it : 1), ckC2 / (it > 0 ? it : 1));
        if (bc.tid == 0 && DART_PHASE_CLOCK_BLOCK)
            printf("phase clocks: it %d A %lld waitAB %lld B %lld waitBC %lld C %lld (per iter A %lld wAB %lld B %lld wBC %lld C %lld)\n", it, ckA, ckW1, ckB, ckW2,
                   ckC, ckA / (it + 1), ckW1 / (it + 1), ckB / (it > 0 ? it : 1), ckW2 / (it > 0 ? it : 1), ckC / (it > 0 ? it : 1));
#endif
#endif
        J = f;
        status = st;
        iters = it;
        kkt = E0;
    }
};

}  // namespace dart
