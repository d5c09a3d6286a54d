// Solver option defaults shared by the C ABI and the host-emulation test harness.
#pragma once
#include <stdlib.h>
#include "solver_core.cuh"
#include "../../include/dart_b200.h"

namespace dart {
inline void fill_opts(const dart_cfg& c, SolverOpts& o) {
    o.tol = c.tol > 0 ? c.tol : 1e-8;
    o.mu0 = c.mu_init > 0 ? c.mu_init : 0.1;
    o.cold = 1;                                            // per launch: dart_solve sets it from warm_w
    o.mu0_auto = c.mu_init > 0 ? 0 : 1;                    // mu_init = 0: the strategy's own default (Solver::start_mu)
    o.kappa_mu = 0.2; o.theta_mu = 1.5; o.kappa_eps = 10.0; o.tau_min = 0.99; o.bound_push = 1e-2;
    o.eta = 1e-4; o.smax = 100.0; o.s_phi = 2.3; o.s_theta = 1.1; o.delta_sw = 1.0;
    o.gamma_theta = 1e-5; o.gamma_phi = 1e-8; o.theta_small = 1e-4;
    o.max_iter = c.max_iter > 0 ? c.max_iter : 200;
    o.max_backtrack = 12;
    o.acc_tol = c.acceptable_tol;
    o.acc_iter = (c.acceptable_iter > 0 && c.acceptable_tol > 0) ? c.acceptable_iter : 0;
    // barrier strategy (IPOPT's mu_strategy), dart_set_barrier_strategy: 2 = DART_BARRIER_AUTO (per method, what measured faster:
    // Model::PC_DEFAULT), 1 = predictor-corrector wherever the kernel has it, 0 = monotone schedule.  The environment
    // variables DART_BARRIER_MONOTONE=1 / DART_BARRIER_MEHROTRA=1 change the default of new handles (A/B runs, host tests).
    o.mehrotra = getenv("DART_BARRIER_MONOTONE") ? 0 : (getenv("DART_BARRIER_MEHROTRA") ? 1 : 2);
}

// Per-launch adjustments of the options (dart_solve, and the host test harness): whether the call carries a warm plan, and the
// one configuration the tiled predictor-corrector step cannot serve -- its corrector recovers the inverse pivots from the
// coupling h_j = H[u_j][carried input j] (Solver::feedforward), which for LMPC is the tilt-rate weight alone (RMPC's rate rows
// always contribute their barrier terms): without a tilt-rate cost LMPC runs the monotone schedule.
inline void launch_opts(const dart_cfg& c, bool has_warm_plan, SolverOpts& o) {
    o.cold = has_warm_plan ? 0 : 1;
    if (c.method == DART_LMPC && (!(c.Rl[2] > 0.0) || !(c.Rl[3] > 0.0))) o.mehrotra = 0;
}
}  // namespace dart
