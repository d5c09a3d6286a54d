// Solver option defaults shared by the C ABI and the host-emulation test harness.
#pragma once
#include <stdlib.h>
#include "solver_core.cuh"
#include "../../include/dart_b200.h"

namespace dart {
inline void fill_opts(const dart_cfg& c, SolverOpts& o) {
    o.tol = c.tol > 0 ? c.tol : 1e-8;
    o.mu0 = c.mu_init > 0 ? c.mu_init : 0.1;
    o.kappa_mu = 0.2; o.theta_mu = 1.5; o.kappa_eps = 10.0; o.tau_min = 0.99; o.bound_push = 1e-2;
    o.eta = 1e-4; o.smax = 100.0; o.s_phi = 2.3; o.s_theta = 1.1; o.delta_sw = 1.0;
    o.gamma_theta = 1e-5; o.gamma_phi = 1e-8; o.theta_small = 1e-4;
    o.max_iter = c.max_iter > 0 ? c.max_iter : 200;
    o.max_backtrack = 12;
    o.acc_tol = c.acceptable_tol;
    o.acc_iter = (c.acceptable_iter > 0 && c.acceptable_tol > 0) ? c.acceptable_iter : 0;
    // barrier strategy (IPOPT's mu_strategy): predictor-corrector where the kernel has it (PMPC axis problems on the scan
    // path), the monotone schedule elsewhere; DART_BARRIER_MONOTONE=1 or dart_set_barrier_strategy select the monotone one
    o.mehrotra = getenv("DART_BARRIER_MONOTONE") ? 0 : 1;
}
}  // namespace dart
