// Plant models, cost weights and constraint rows of the three reference controllers, in the form
// the generic solver (solver_core.cuh) consumes.  Dynamics are the reference's RK4 steps with analytic
// forward sensitivities (A = dF/dx, B = dF/du) propagated through the four stages.
//
//   PmpcAxis  PMPC/src/controller/mpc_3d.py:87-104 (dynamics), :28-85 (NLP)
//   Rmpc      RMPC/dev_dual/controller/np_mpc_adaptive_with_linear_regressor.py:171-193, :65-168
//   LmpcAxis  LMPC/src/controller/rlmpc2.py:260-436, :438-491
//
// Structure used (verified against the full coupled NLPs by the oracle, which does not use it):
//  * PMPC's x and y motions are independent 2-state/1-input problems with a separable cost; the z rows
//    (mpc_3d.py:93-95) have no cost, no bound and feed nothing back, so they are rolled out at the end.
//  * LMPC's 8-state model splits into {px,vx,theta_y,omega_y | a} and {py,vy,theta_x,omega_x | b}
//    (rlmpc2.py:366-410: v_slip_x couples vx with omega_y, tau_topple_y uses m_x, and symmetrically).
//  * RMPC couples x and y through the 14 regressor parameters and is solved as one problem.
// Tilt-rate terms (u_k - u_{k-1}) carry the previous input as NU extra trailing states.
#pragma once
#include "solver_core.cuh"
#include "../../include/dart_b200.h"

#ifndef DART_ROLL_RK4
#define DART_ROLL_RK4 1
#endif
#ifndef DART_FOLD_RK4
#define DART_FOLD_RK4 0
#endif

namespace dart {

struct KArgs {
    int B, N;
    SolverOpts o;
    dart_cfg cfg;
    const double *x0, *ref, *aux, *warm;
    double *w_out, *u0, *J;
    int32_t *status, *iters;
    double* rows;      // optional packed result rows [B,4] = [u0x, u0y, J, status] for the multi-GPU gather
    double* dual;      // optional dual-state blocks [B, NAXIS * dual_doubles(N)] (read when warm is given, always written)
    // optional hand-over area for models whose axes are solved by different warps: per-axis [J, status, iters, kkt]
    // records [B, NAXIS, 4] and an arrival counter per instance (zero between launches); see solve_block
    double* axis_part;
    int32_t* axis_sync;
    // optional peer result rows (multi-GPU gather without a collective): every instance's row [u0x, u0y, J, status] is ALSO
    // stored at peer_rows[p] + (peer_off + inst) * 4 for p < n_peers -- peer-mapped gathered buffers of the other GPUs (and
    // this GPU's own), written over NVLink from the solve kernel's epilogue
    double* peer_rows[DART_MAX_PEERS];
    int32_t n_peers;
    int64_t peer_off;
};

// the packed result row of one instance: local rows buffer and / or every peer's gathered buffer
DART_HD void store_result_row(const KArgs& a, int inst, double u0x, double u0y, double J, double st) {
    if (a.rows) {
        double* r = a.rows + (long)inst * 4;
        r[0] = u0x; r[1] = u0y; r[2] = J; r[3] = st;
    }
    for (int p = 0; p < a.n_peers; ++p) {
        double* r = a.peer_rows[p] + (a.peer_off + inst) * 4;
        r[0] = u0x; r[1] = u0y; r[2] = J; r[3] = st;
    }
}

// Generic RK4 step with forward sensitivities.  Md::deriv(prm, x, u, f, fx, fu) gives xdot and its
// Jacobians for the NP physical states.
template <class Md>
DART_HD void rk4_sens(const typename Md::Prm& p, const double* x, const double* u, double h, double* F,
                      double* A, double* Bm, double* tanu) {
    constexpr int np = Md::NP, m = Md::NU;
    double k[np], fx[np * np], fu[np * m], Sx[np * np], Su[np * m], xs[np], acc[np];
    // the input is held over the step and enters only through sin/cos: evaluate them once
    double su[m], cu[m];
    DART_UNROLL for (int j = 0; j < m; ++j) { sincos(u[j], &su[j], &cu[j]); tanu[j] = su[j] / cu[j]; }
    // Md::fx_kind(a, q) / fu_kind(a, j): 0 = general entry, 1 = structurally zero, 2 = exactly one.  The kinematic rows
    // (position' = velocity) are unit rows: their products drop out at compile time (IEEE arithmetic does not let the
    // compiler fold 0 * x or 1 * x itself) -- same values, 40-50 % fewer FMAs in the sensitivity propagation.
    auto fxv = [&](int a, int q) { return Md::fx_kind(a, q) == 0 ? fx[a * np + q] : (Md::fx_kind(a, q) == 2 ? 1.0 : 0.0); };
    auto fuv = [&](int a, int j) { return Md::fu_kind(a, j) == 0 ? fu[a * m + j] : (Md::fu_kind(a, j) == 2 ? 1.0 : 0.0); };
    // DART_ROLL_RK4 (models with np > 2): stages 2-4 stay a rolled LOOP.  The stage body (derivative with its
    // transcendentals + sensitivity propagation) is the bulk of the solve kernel's code and every eval1 call site carries a
    // copy of it; the LMPC kernel's largest stall reason was instruction fetch (no_instruction 30 % of the stall samples;
    // instruction cache 32 kB, kernel 180 kB).  Measured on one box: LMPC 6.50 -> 6.07 ms, RMPC 1.47 -> 1.41 ms; the small
    // PMPC body is faster unrolled.  DART_FOLD_RK4 also runs stage 1 through the loop (as the general stage with S = 0, c = 0:
    // the same values bit for bit, another 600 instructions less code) -- no measurable difference in a same-box A/B (LMPC
    // 6.00 vs 6.02 ms: the extra FMAs on zeros cost what the fetches save), so it is off.
    constexpr bool kRoll = DART_ROLL_RK4 && Md::NP > 2;
    constexpr bool kFold = DART_FOLD_RK4 && kRoll;
    if (kFold) {
        DART_UNROLL for (int i = 0; i < np * np; ++i) { Sx[i] = 0.0; A[i] = 0.0; }
        DART_UNROLL for (int i = 0; i < np * m; ++i) { Su[i] = 0.0; Bm[i] = 0.0; }
        DART_UNROLL for (int i = 0; i < np; ++i) { acc[i] = 0.0; xs[i] = x[i]; }
    } else {
        // stage 1
        Md::deriv(p, x, su, cu, k, fx, fu);
        DART_UNROLL for (int a = 0; a < np; ++a) {
            DART_UNROLL for (int b = 0; b < np; ++b) { Sx[a * np + b] = fxv(a, b); A[a * np + b] = fxv(a, b); }
            DART_UNROLL for (int j = 0; j < m; ++j) { Su[a * m + j] = fuv(a, j); Bm[a * m + j] = fuv(a, j); }
        }
        DART_UNROLL for (int i = 0; i < np; ++i) { acc[i] = k[i]; xs[i] = x[i] + 0.5 * h * k[i]; }
    }
    DART_UNROLL_N(kRoll ? 1 : 3)
    for (int st = kFold ? 1 : 2; st <= 4; ++st) {
        const double c = (st == 4) ? h : (st == 1 ? 0.0 : 0.5 * h);      // step used to reach this stage's evaluation point
        const double wgt = (st == 4 || st == 1) ? 1.0 : 2.0;
        Md::deriv(p, xs, su, cu, k, fx, fu);
        double Nx[np * np], Nu[np * m];
        DART_UNROLL for (int a = 0; a < np; ++a) {
            DART_UNROLL for (int b = 0; b < np; ++b) {
                double v = fxv(a, b);
                DART_UNROLL for (int q = 0; q < np; ++q) {
                    if (Md::fx_kind(a, q) == 1) continue;
                    if (Md::fx_kind(a, q) == 2) v += c * Sx[q * np + b];
                    else v += fx[a * np + q] * (c * Sx[q * np + b]);
                }
                Nx[a * np + b] = v;
            }
            DART_UNROLL for (int j = 0; j < m; ++j) {
                double v = fuv(a, j);
                DART_UNROLL for (int q = 0; q < np; ++q) {
                    if (Md::fx_kind(a, q) == 1) continue;
                    if (Md::fx_kind(a, q) == 2) v += c * Su[q * m + j];
                    else v += fx[a * np + q] * (c * Su[q * m + j]);
                }
                Nu[a * m + j] = v;
            }
        }
        DART_UNROLL for (int i = 0; i < np * np; ++i) { Sx[i] = Nx[i]; A[i] += wgt * Nx[i]; }
        DART_UNROLL for (int i = 0; i < np * m; ++i) { Su[i] = Nu[i]; Bm[i] += wgt * Nu[i]; }
        const double cn = (st == 3) ? h : 0.5 * h;     // step to the next stage's evaluation point
        DART_UNROLL for (int i = 0; i < np; ++i) { acc[i] += wgt * k[i]; xs[i] = x[i] + cn * k[i]; }
    }
    DART_UNROLL for (int i = 0; i < np; ++i) F[i] = x[i] + h / 6.0 * acc[i];
    DART_UNROLL for (int a = 0; a < np; ++a)
        DART_UNROLL for (int b = 0; b < np; ++b) A[a * np + b] = (a == b ? 1.0 : 0.0) + h / 6.0 * A[a * np + b];
    DART_UNROLL for (int i = 0; i < np * m; ++i) Bm[i] = h / 6.0 * Bm[i];
}

// =============================================================================================== PMPC
struct PmpcAxis {
    static constexpr int NX = 2, NU = 1, NR = 1, NP = 2, NAUG = 0, NAXIS = 2;
    static constexpr bool MEHROTRA = true;      // Solver::sweeps_scan_pc (solver_core.cuh)
    static constexpr bool PC_COLD = true;
    static constexpr bool PC_DEFAULT = true;    // DART_BARRIER_AUTO: 9.0 -> 5.7 iterations, 0.082 -> 0.068 ms at the headline batch
    // cold-start multiplier scale under predictor-corrector steps (Solver::start_mu): 0.01 cuts the MEAN (5.7 -> 5.1 iterations,
    // filled GPU 3.78 -> 3.70 ms) but not the slowest instance, which is what the headline launch waits for (0.0677 -> 0.0710 ms)
    static constexpr double MU0_PC = 0.1;
    static constexpr bool SERIAL_RICCATI = true;
    // structure the serial sweep may rely on: the position does not enter the dynamics, so column 0 of the RK4
    // sensitivity A is exactly e_0
    DART_HD static constexpr int a_kind(int a, int b) { return b == 0 ? (a == 0 ? 2 : 1) : 0; }
    // structure of the continuous-time Jacobians (deriv below): fx = [[0, 1], [0, -mu]], fu = [0, g cos u]
    DART_HD static constexpr int fx_kind(int a, int q) { return a == 0 ? (q == 1 ? 2 : 1) : (q == 0 ? 1 : 0); }
    DART_HD static constexpr int fu_kind(int a, int) { return a == 0 ? 1 : 0; }
    static constexpr int MAX_THREADS = 256, MIN_BLOCKS = 1, BT_LARGE = 32;   // BT_LARGE: block size when the GPU is filled (measured: 32 < 64 < 128)
    static constexpr int NXF = 6;   // states per stage in the reference's decision vector
    static constexpr int NDEF = 15; // the reference's horizon (compile-time instantiation)
    struct Prm { double Qp, Qv, R, mu, g, Ts, ulo, uhi, rp, rv; };

    DART_HD static int ref_doubles(int) { return 0; }
    DART_HD static int nw(int N) { return (N + 1) * 6 + N * 2; }
    DART_HD static int nx_in() { return 6; }
    DART_HD static int nref_in(int) { return 6; }
    DART_HD static int naux_in() { return 4; }

    DART_HD static void deriv(const Prm& p, const double* x, const double* su, const double* cu, double* f, double* fx,
                              double* fu) {
        f[0] = x[1];
        f[1] = p.g * su[0] - p.mu * x[1];
        fx[0] = 0.0; fx[1] = 1.0; fx[2] = 0.0; fx[3] = -p.mu;
        fu[0] = 0.0; fu[1] = p.g * cu[0];
    }
    DART_HD static void dyn(const Prm& p, const double* x, const double* u, double* F, double* A, double* Bm, double* tanu) {
        rk4_sens<PmpcAxis>(p, x, u, p.Ts, F, A, Bm, tanu);
    }
    DART_HD static double wy(const Prm& p, int i) { return i == 0 ? p.Qp : (i == 1 ? p.Qv : p.R); }
    DART_HD static double ry(const Prm& p, const double*, int, int i) { return i == 0 ? p.rp : (i == 1 ? p.rv : 0.0); }
    DART_HD static double wd(const Prm&, int) { return 0.0; }
    DART_HD static double wT(const Prm& p, int i) { return i == 0 ? p.Qp : p.Qv; }
    DART_HD static double rT(const Prm& p, const double*, int, int i) { return i == 0 ? p.rp : p.rv; }
    DART_HD static constexpr int row_ia(int) { return 2; }
    DART_HD static constexpr int row_ib(int) { return -1; }
    DART_HD static constexpr double row_sa(int) { return 1.0; }
    DART_HD static constexpr double row_sb(int) { return 0.0; }
    DART_HD static constexpr bool row_skip0(int) { return false; }
    DART_HD static void bounds(const Prm& p, int, double& lo, double& hi) { lo = p.ulo; hi = p.uhi; }

    DART_HD static void load(Prm& p, const KArgs& a, int inst, int axis) {
        const dart_cfg& c = a.cfg;
        if (a.aux) {
            const double* q = a.aux + (long)inst * 4;
            p.Qp = q[0]; p.Qv = q[1]; p.R = q[2]; p.mu = q[3];
        } else {
            p.Qp = c.Qp; p.Qv = c.Qv; p.R = c.R; p.mu = c.mu;
        }
        p.g = c.g; p.Ts = c.Ts; p.ulo = c.u_lo; p.uhi = c.u_hi;
        p.rp = a.ref[(long)inst * 6 + 2 * axis];
        p.rv = a.ref[(long)inst * 6 + 2 * axis + 1];
    }
    DART_HD static bool infeasible0(const Prm&, const double*) { return false; }
    // index of sub-problem state i inside the reference's per-stage state vector
    DART_HD static int xmap(int axis, int i) { return 2 * axis + i; }
    DART_HD static void x0(const KArgs& a, int inst, int axis, double* x) {
        x[0] = a.x0[(long)inst * 6 + 2 * axis];
        x[1] = a.x0[(long)inst * 6 + 2 * axis + 1];
    }
    template <class T>
    DART_HD static void load_ref(const T&, const KArgs&, int, double*) {}
};

// =============================================================================================== RMPC
struct Rmpc {
    static constexpr int NX = 6, NU = 2, NR = 6, NP = 4, NAUG = 2, NAXIS = 1;
    static constexpr bool MEHROTRA = true;       // Solver::pc_rows + corrector_tile (solver_core.cuh)
    // DART_BARRIER_AUTO: predictor-corrector steps for cold-started launches (12.95 -> 8.85 iterations, each 1.31x dearer:
    // 1.432 -> 1.334 ms at 4096 instances), the monotone schedule for warm-started ones -- the closed loop's warm-start layer
    // (mu_init 1e-4 + previous plan: 5.5 iterations per solve) gains nothing from an adaptive mu and would pay the dearer iterations
    static constexpr bool PC_DEFAULT = false, PC_COLD = true;
    static constexpr double MU0_PC = 0.01;       // 9.43 -> 8.85 iterations, 1.375 -> 1.334 ms (predictor-corrector, cold start)
    static constexpr bool SERIAL_RICCATI = false;
    DART_HD static constexpr int a_kind(int, int) { return 0; }
    // rows 0 and 2 are the kinematic unit rows (p' = v); rows 1 and 3 are dense; u_j enters the acceleration of axis j only
    DART_HD static constexpr int fx_kind(int a, int q) { return (a == 0 || a == 2) ? (q == a + 1 ? 2 : 1) : 0; }
    DART_HD static constexpr int fu_kind(int a, int j) { return ((a == 1 && j == 0) || (a == 3 && j == 1)) ? 0 : 1; }
#ifndef DART_RMPC_MAXT
#define DART_RMPC_MAXT 256
#endif
#ifndef DART_RMPC_MINB
#define DART_RMPC_MINB 1
#endif
    static constexpr int MAX_THREADS = DART_RMPC_MAXT, MIN_BLOCKS = DART_RMPC_MINB, BT_LARGE = 32;
    static constexpr int NXF = 4;
    static constexpr int NDEF = 20;
    struct Prm { double Qp, Qv, Ru, Rdu, gz, Ts, ulo, uhi, dlo, dhi, vmax, inv_eps; double th[14]; };

    DART_HD static int ref_doubles(int N) { return (N + 1) * 4; }
    DART_HD static int nw(int N) { return (N + 1) * 4 + N * 2; }
    DART_HD static int nx_in() { return 4; }
    DART_HD static int nref_in(int N) { return (N + 1) * 4; }
    DART_HD static int naux_in() { return 16; }

    DART_HD static void deriv(const Prm& p, const double* x, const double* su, const double* cu, double* f, double* fx,
                              double* fu) {
        const double sa = su[0], ca = cu[0], sb = su[1], cb = cu[1];
        const double tx = tanh(x[1] * p.inv_eps), ty = tanh(x[3] * p.inv_eps);
        const double dtx = (1.0 - tx * tx) * p.inv_eps, dty = (1.0 - ty * ty) * p.inv_eps;
        const double* a = p.th;
        const double* b = p.th + 7;
        f[0] = x[1];
        f[1] = p.gz * sa + (a[0] * x[0] + a[1] * x[1] + a[2] * x[2] + a[3] * x[3] + a[4] * tx + a[5] * ty + a[6]);
        f[2] = x[3];
        f[3] = p.gz * sb + (b[0] * x[0] + b[1] * x[1] + b[2] * x[2] + b[3] * x[3] + b[4] * tx + b[5] * ty + b[6]);
        DART_UNROLL for (int i = 0; i < 16; ++i) fx[i] = 0.0;
        fx[0 * 4 + 1] = 1.0;
        fx[1 * 4 + 0] = a[0]; fx[1 * 4 + 1] = a[1] + a[4] * dtx; fx[1 * 4 + 2] = a[2]; fx[1 * 4 + 3] = a[3] + a[5] * dty;
        fx[2 * 4 + 3] = 1.0;
        fx[3 * 4 + 0] = b[0]; fx[3 * 4 + 1] = b[1] + b[4] * dtx; fx[3 * 4 + 2] = b[2]; fx[3 * 4 + 3] = b[3] + b[5] * dty;
        DART_UNROLL for (int i = 0; i < 8; ++i) fu[i] = 0.0;
        fu[1 * 2 + 0] = p.gz * ca;
        fu[3 * 2 + 1] = p.gz * cb;
    }
    DART_HD static void dyn(const Prm& p, const double* x, const double* u, double* F, double* A, double* Bm, double* tanu) {
        rk4_sens<Rmpc>(p, x, u, p.Ts, F, A, Bm, tanu);
    }
    DART_HD static double wy(const Prm& p, int i) {
        return (i == 0 || i == 2) ? p.Qp : ((i == 1 || i == 3) ? p.Qv : (i >= 6 ? p.Ru : 0.0));
    }
    DART_HD static double ry(const Prm&, const double* REF, int k, int i) { return i < 4 ? REF[k * 4 + i] : 0.0; }
    DART_HD static double wd(const Prm& p, int) { return p.Rdu; }
    DART_HD static double wT(const Prm& p, int i) { return (i == 0 || i == 2) ? p.Qp : ((i == 1 || i == 3) ? p.Qv : 0.0); }
    DART_HD static double rT(const Prm&, const double* REF, int N, int i) { return i < 4 ? REF[N * 4 + i] : 0.0; }
    // rows: u0, u1 boxes | du0 = u0 - uprev0, du1 | vx, vy caps (act on fixed x0 at k = 0)
    DART_HD static constexpr int row_ia(int r) { return r == 0 ? 6 : r == 1 ? 7 : r == 2 ? 6 : r == 3 ? 7 : r == 4 ? 1 : 3; }
    DART_HD static constexpr int row_ib(int r) { return r == 2 ? 4 : r == 3 ? 5 : -1; }
    DART_HD static constexpr double row_sa(int) { return 1.0; }
    DART_HD static constexpr double row_sb(int r) { return (r == 2 || r == 3) ? -1.0 : 0.0; }
    DART_HD static constexpr bool row_skip0(int r) { return r >= 4; }
    DART_HD static void bounds(const Prm& p, int r, double& lo, double& hi) {
        if (r < 2) { lo = p.ulo; hi = p.uhi; }
        else if (r < 4) { lo = p.dlo; hi = p.dhi; }
        else { lo = -p.vmax; hi = p.vmax; }
    }
    DART_HD static void load(Prm& p, const KArgs& a, int inst, int) {
        const dart_cfg& c = a.cfg;
        p.Qp = c.Qp; p.Qv = c.Qv; p.Ru = c.R; p.Rdu = c.Rdu; p.gz = c.g; p.Ts = c.Ts;
        p.ulo = c.u_lo; p.uhi = c.u_hi; p.dlo = c.du_lo; p.dhi = c.du_hi; p.vmax = c.vmax; p.inv_eps = 1.0 / c.v_eps;
        const double* q = a.aux + (long)inst * 16;
        DART_UNROLL for (int i = 0; i < 14; ++i) p.th[i] = q[2 + i];
    }
    DART_HD static bool infeasible0(const Prm& p, const double* x) { return fabs(x[1]) > p.vmax || fabs(x[3]) > p.vmax; }
    DART_HD static int xmap(int, int i) { return i; }
    DART_HD static void x0(const KArgs& a, int inst, int, double* x) {
        DART_UNROLL for (int i = 0; i < 4; ++i) x[i] = a.x0[(long)inst * 4 + i];
        x[4] = a.aux[(long)inst * 16 + 0];
        x[5] = a.aux[(long)inst * 16 + 1];
    }
    template <class T>
    DART_HD static void load_ref(const T& tile, const KArgs& a, int inst, double* REF) {
        const int cnt = (a.N + 1) * 4;
        for (int i = tile.lane(); i < cnt; i += tile.size()) REF[i] = a.ref[(long)inst * cnt + i];
    }
};

// =============================================================================================== LMPC
struct LmpcAxis {
    static constexpr int NX = 5, NU = 1, NR = 1, NP = 4, NAUG = 1, NAXIS = 2;
    static constexpr bool MEHROTRA = true;       // Solver::pc_rows + corrector_tile (solver_core.cuh)
    static constexpr bool PC_COLD = true;
    static constexpr bool PC_DEFAULT = true;     // DART_BARRIER_AUTO: 10.8 -> 7.2 iterations, each 29 % dearer: 5.99 -> 5.18 ms at 16 384
    static constexpr double MU0_PC = 0.01;       // 7.25 -> 5.89 iterations, 5.05 -> 4.22 ms at 16 384 (cold start)
    static constexpr bool SERIAL_RICCATI = false;   // measured: the tiled sweep is 1.3-2x faster than the per-lane one at n = 5
    DART_HD static constexpr int a_kind(int, int) { return 0; }
    // rows 0 and 2 kinematic; row 1 (translation) does not see the angle, row 3 (rotation) does not see the position
    DART_HD static constexpr int fx_kind(int a, int q) {
        return (a == 0 || a == 2) ? (q == a + 1 ? 2 : 1) : (a == 1 ? (q == 2 ? 1 : 0) : (q == 0 ? 1 : 0));
    }
    DART_HD static constexpr int fu_kind(int a, int) { return a == 1 ? 0 : 1; }
    // one instance (two axis tiles) per block; 168 registers -> 6 blocks = 12 warps per SM (shared memory allows 14).
    // Measured at 16 384 instances: bounds (64,5) 7.06 ms, (128,3) 7.58 ms (same register count, worse schedule),
    // (64,4) / (128,2) with 255 registers and 8 warps 7.30 ms, (64,7) with 128 registers 8.2 ms
#ifndef DART_LMPC_MAXT
#define DART_LMPC_MAXT 64
#endif
#ifndef DART_LMPC_MINB
#define DART_LMPC_MINB 5
#endif
    static constexpr int MAX_THREADS = DART_LMPC_MAXT, MIN_BLOCKS = DART_LMPC_MINB, BT_LARGE = 64;
    static constexpr int NXF = 8;
    static constexpr int NDEF = 20;
    struct Prm {
        double Ts, ulo, uhi;
        double Q[4], Qt[4], Ru, Rdu, ref[4];
        // translational: m, c, k, Stribeck (Fs, Fc, Bv, vs, eps); rolling radius r (torque) and rs (slip sign)
        double m, c, k, Fs, Fc, Bv, ivs, ieps, r, rs;
        // rotational: inertia (+1e-12), damping, Stribeck, com height
        double iI, crot, Fsr, Fcr, Br, ivsr, iepsr, mgh;
    };
    DART_HD static int ref_doubles(int) { return 0; }
    DART_HD static int nw(int N) { return (N + 1) * 8 + N * 2; }
    DART_HD static int nx_in() { return 8; }
    DART_HD static int nref_in(int) { return 8; }
    DART_HD static int naux_in() { return 36; }

    // stribeck_fric (rlmpc2.py:355-359) and its derivative; d|v|/dv = sign(v), sign(0) = 0 (CasADi fabs)
    DART_HD static void stribeck(double v, double Fs, double Fc, double B, double ivs, double ieps, double& S, double& dS) {
        const double av = fabs(v);
        const double sg = (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0);
        const double E = exp(-av * ivs);
        const double t = tanh(v * ieps);
        const double lvl = Fc + (Fs - Fc) * E;
        S = t * lvl + B * v;
        dS = (1.0 - t * t) * ieps * lvl + t * (Fs - Fc) * E * (-sg * ivs) + B;
    }
    DART_HD static void deriv(const Prm& p, const double* x, const double* su, const double* cu, double* f, double* fx,
                              double* fu) {
        const double pos = x[0], v = x[1], th = x[2], om = x[3];
        const double sa = su[0], ca = cu[0];
        double st, ct;
        sincos(th, &st, &ct);
        double Ff, dFf, Fr, dFr, Tr, dTr;
        stribeck(v, p.Fs, p.Fc, p.Bv, p.ivs, p.ieps, Ff, dFf);
        const double vslip = v - p.rs * om;
        stribeck(vslip, p.Fs, p.Fc, p.Bv, p.ivs, p.ieps, Fr, dFr);
        stribeck(om, p.Fsr, p.Fcr, p.Br, p.ivsr, p.iepsr, Tr, dTr);
        const double im = 1.0 / p.m;
        f[0] = v;
        f[1] = (p.m * (9.81 * sa) - p.c * v - p.k * pos - Ff - Fr) * im;
        f[2] = om;
        f[3] = (-p.r * Fr - Tr - p.crot * om - p.mgh * st) * p.iI;
        DART_UNROLL for (int i = 0; i < 16; ++i) fx[i] = 0.0;
        fx[0 * 4 + 1] = 1.0;
        fx[1 * 4 + 0] = -p.k * im;
        fx[1 * 4 + 1] = (-p.c - dFf - dFr) * im;
        fx[1 * 4 + 3] = (p.rs * dFr) * im;
        fx[2 * 4 + 3] = 1.0;
        fx[3 * 4 + 1] = (-p.r * dFr) * p.iI;
        fx[3 * 4 + 2] = (-p.mgh * ct) * p.iI;
        fx[3 * 4 + 3] = (p.r * p.rs * dFr - dTr - p.crot) * p.iI;
        fu[0] = 0.0; fu[1] = 9.81 * ca; fu[2] = 0.0; fu[3] = 0.0;
    }
    DART_HD static void dyn(const Prm& p, const double* x, const double* u, double* F, double* A, double* Bm, double* tanu) {
        rk4_sens<LmpcAxis>(p, x, u, p.Ts, F, A, Bm, tanu);
    }
    DART_HD static double wy(const Prm& p, int i) { return i < 4 ? p.Q[i] : (i == 4 ? 0.0 : p.Ru); }
    DART_HD static double ry(const Prm& p, const double*, int, int i) { return i < 4 ? p.ref[i] : 0.0; }
    DART_HD static double wd(const Prm& p, int) { return p.Rdu; }
    DART_HD static double wT(const Prm& p, int i) { return i < 4 ? p.Qt[i] : 0.0; }
    DART_HD static double rT(const Prm& p, const double*, int, int i) { return i < 4 ? p.ref[i] : 0.0; }
    DART_HD static constexpr int row_ia(int) { return 5; }
    DART_HD static constexpr int row_ib(int) { return -1; }
    DART_HD static constexpr double row_sa(int) { return 1.0; }
    DART_HD static constexpr double row_sb(int) { return 0.0; }
    DART_HD static constexpr bool row_skip0(int) { return false; }
    DART_HD static void bounds(const Prm& p, int, double& lo, double& hi) { lo = p.ulo; hi = p.uhi; }

    DART_HD static double sq(double v) { return fabs(v) + 1e-6; }   // squash_param, rlmpc2.py:287-289
    // full-state index of sub-problem state i: axis 0 = {px, vx, theta_y, omega_y}, axis 1 = {py, vy, theta_x, omega_x}
    DART_HD static int xmap(int axis, int i) { return axis == 0 ? (i < 2 ? i : 4 + i) : (i < 2 ? 2 + i : 2 + i); }
    DART_HD static void load(Prm& p, const KArgs& a, int inst, int axis) {
        const dart_cfg& c = a.cfg;
        const double* pv = a.aux + (long)inst * 36 + 2;
        p.Ts = c.Ts; p.ulo = c.u_lo; p.uhi = c.u_hi;
        DART_UNROLL for (int i = 0; i < 4; ++i) {
            const int gi = xmap(axis, i);
            p.Q[i] = c.Q[gi]; p.Qt[i] = c.Qt[gi]; p.ref[i] = a.ref[(long)inst * 8 + gi];
        }
        p.Ru = c.Rl[axis]; p.Rdu = c.Rl[2 + axis];
        if (axis == 0) {
            p.m = sq(pv[0]); p.c = sq(pv[2]); p.k = sq(pv[4]);
            p.Fs = pv[6]; p.Fc = pv[7]; p.Bv = pv[8]; p.ivs = 1.0 / (sq(pv[9]) + 1e-12); p.ieps = 1.0 / sq(pv[10]);
            p.r = sq(pv[18]); p.rs = p.r;
            p.iI = 1.0 / (sq(pv[17]) + 1e-12); p.crot = sq(pv[21]);
            p.Fsr = pv[27]; p.Fcr = pv[28]; p.Br = pv[29]; p.ivsr = 1.0 / (sq(pv[30]) + 1e-12); p.iepsr = 1.0 / sq(pv[31]);
            p.mgh = p.m * 9.81 * sq(pv[33]);
        } else {
            p.m = sq(pv[1]); p.c = sq(pv[3]); p.k = sq(pv[5]);
            p.Fs = pv[11]; p.Fc = pv[12]; p.Bv = pv[13]; p.ivs = 1.0 / (sq(pv[14]) + 1e-12); p.ieps = 1.0 / sq(pv[15]);
            p.r = sq(pv[19]); p.rs = -p.r;
            p.iI = 1.0 / (sq(pv[16]) + 1e-12); p.crot = sq(pv[20]);
            p.Fsr = pv[22]; p.Fcr = pv[23]; p.Br = pv[24]; p.ivsr = 1.0 / (sq(pv[25]) + 1e-12); p.iepsr = 1.0 / sq(pv[26]);
            p.mgh = p.m * 9.81 * sq(pv[32]);
        }
    }
    DART_HD static bool infeasible0(const Prm&, const double*) { return false; }
    DART_HD static void x0(const KArgs& a, int inst, int axis, double* x) {
        DART_UNROLL for (int i = 0; i < 4; ++i) x[i] = a.x0[(long)inst * 8 + xmap(axis, i)];
        x[4] = a.aux[(long)inst * 36 + axis];
    }
    template <class T>
    DART_HD static void load_ref(const T&, const KArgs&, int, double*) {}
};

// =============================================================================================== driver
// Solve sub-problem (inst, axis) with the lanes of `tile`; `base` is this tile's workspace.
// Returns per-problem J/status/iters/kkt; writes X/U of this sub-problem into w_out (reference layout).
template <class M, class T, int NC = 0>
DART_HD void solve_one(const T& tile, const KArgs& a, int inst, int axis, bool active, bool write, const BlockCtx& bc, double* base,
                       double& J, int32_t& status, int32_t& iters, double& kkt) {
    constexpr int n = M::NX, m = M::NU, np = M::NP;
    const int N = (NC > 0) ? NC : a.N;
    Workspace<M> w;
    w.bind(base, N);
    typename M::Prm prm;
    double x0[n];
    const int nwf = M::nw(N);
    const int uoff = (N + 1) * M::NXF;
    const int ucol = (M::NAXIS > 1) ? axis : 0;
    // the starting point: the caller's plan (warm) or x0 held over the horizon with zero inputs (mpc_3d.py:123)
    auto load_start = [&]() {
        const double* warm = a.warm ? a.warm + (long)inst * nwf : nullptr;     // re-derived, not kept live across the solve
        for (int k = tile.lane(); k <= N; k += tile.size()) {
            DART_UNROLL for (int i = 0; i < np; ++i)
                w.X[k * n + i] = (warm && k > 0) ? warm[k * M::NXF + M::xmap(axis, i)] : x0[i];
            // carried previous input: u_prev at k = 0, U[k-1] afterwards (so tilt-rate rows start consistent)
            DART_UNROLL for (int i = np; i < n; ++i)
                w.X[k * n + i] = (k == 0) ? x0[i] : (warm ? warm[uoff + (k - 1) * 2 + (M::NAXIS > 1 ? ucol : (i - np))] : 0.0);
            if (k < N) {
                DART_UNROLL for (int j = 0; j < m; ++j)
                    w.U[k * m + j] = warm ? warm[uoff + k * 2 + (M::NAXIS > 1 ? ucol : j)] : 0.0;
            }
        }
        tile.sync();
    };
    if (active) {
        M::load(prm, a, inst, axis);
        M::x0(a, inst, axis, x0);
        M::load_ref(tile, a, inst, w.REF);
        load_start();
    }
    // every tile of the block (active or not) takes part in run(): its threads also serve the serial-sweep phase
    Solver<M, T, NC> s(tile, prm, a.o, N, w, bc);
    // dual warm start (dart_set_dual_state): the block pointer is re-derived where it is needed so that it does not
    // occupy registers across the solve
    auto dual_block = [&]() { return a.dual + ((long)inst * M::NAXIS + axis) * Solver<M, T, NC>::dual_doubles(N); };
    double mu0 = Solver<M, T, NC>::start_mu(a.o);
    if (active) {
        const bool dualwarm = a.dual != nullptr && a.warm != nullptr && dual_block()[0] == 1.0;
        if (a.dual != nullptr && !dualwarm) mu0 = dmax(mu0, 1e-4);     // no usable dual state: never below the primal-warm value
        s.init_rows(mu0);
        if (T::kLockstep) {
            // lockstep tiles: the sibling tile of the warp may have a usable dual block when this one has not
            if (a.dual != nullptr && a.warm != nullptr) s.load_duals(dual_block(), mu0, dualwarm);
        } else if (dualwarm) {
            s.load_duals(dual_block(), mu0);
        }
    }
    s.run(active, mu0, J, status, iters, kkt);
    if (write && a.dual != nullptr) s.store_duals(dual_block(), status == ST_CONVERGED || status == ST_ACCEPTABLE);
    if (!write) return;
    if (M::infeasible0(prm, x0) && status != ST_NUMERIC) status = ST_INFEASIBLE;
    // a solve that ran into NaN/Inf hands back its starting point, not the broken iterate: a closed loop that feeds
    // plans and commands back in must not be poisoned by one failed solve (selected at the output, so that the cold
    // path adds nothing to the solve's register budget)
    const bool bad = status == ST_NUMERIC;
    const double* warm = a.warm ? a.warm + (long)inst * nwf : nullptr;
    if (a.w_out) {
        double* wo = a.w_out + (long)inst * nwf;
        for (int k = tile.lane(); k <= N; k += tile.size()) {
            DART_UNROLL for (int i = 0; i < np; ++i) {
                const int io = k * M::NXF + M::xmap(axis, i);
                wo[io] = bad ? ((warm && k > 0) ? warm[io] : x0[i]) : w.X[k * n + i];
            }
            if (k < N) {
                DART_UNROLL for (int j = 0; j < m; ++j) {
                    const int io = uoff + k * 2 + (M::NAXIS > 1 ? ucol : j);
                    wo[io] = bad ? (warm ? warm[io] : 0.0) : w.U[k * m + j];
                }
            }
        }
    }
    if (tile.lane() == 0) {
        DART_UNROLL for (int j = 0; j < m; ++j) {
            const int jo = (M::NAXIS > 1 ? ucol : j);
            a.u0[(long)inst * 2 + jo] = bad ? (warm ? warm[uoff + jo] : 0.0) : w.U[j];
        }
    }
}

// PMPC z rows (mpc_3d.py:93-95 through _rk4_step :99-104): literal RK4 of [pz, vz] given both tilts.
DART_HD void pmpc_z_rollout(const KArgs& a, int inst) {
    if (!a.w_out) return;
    const int N = a.N;
    double* wo = a.w_out + (long)inst * ((N + 1) * 6 + N * 2);
    const double g = a.cfg.g, Ts = a.cfg.Ts;
    double pz = a.x0[(long)inst * 6 + 4], vz = a.x0[(long)inst * 6 + 5];
    wo[4] = pz; wo[5] = vz;
    const double* U = wo + (N + 1) * 6;
    for (int k = 0; k < N; ++k) {
        const double tx = U[k * 2], ty = U[k * 2 + 1];
        const double vn = -g * (tx * tx + ty * ty);
        // k_i = [vn, (vn - vz_i)/Ts]
        const double a1 = (vn - vz) / Ts;
        const double a2 = (vn - (vz + Ts / 2 * a1)) / Ts;
        const double a3 = (vn - (vz + Ts / 2 * a2)) / Ts;
        const double a4 = (vn - (vz + Ts * a3)) / Ts;
        pz = pz + Ts / 6 * (vn + 2 * vn + 2 * vn + vn);
        vz = vz + Ts / 6 * (a1 + 2 * a2 + 2 * a3 + a4);
        wo[(k + 1) * 6 + 4] = pz;
        wo[(k + 1) * 6 + 5] = vz;
    }
}

}  // namespace dart
