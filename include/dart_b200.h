/* dart_b200 -- C ABI of the B200-native batched tray-tilt NMPC engine.
 *
 * Drop-in boundary for the high-level MPC solve of DART (dart-icra/DART-Dual-Arm-Non-Prehensile-
 * Manipulation).  The reference has no FFI: its boundary is three Python classes and one queue
 * protocol (citations relative to the reference tree):
 *
 *   PMPC.solve(target)                       PMPC/src/controller/mpc_3d.py:115-138
 *   mpc_worker(...) queue service loop       PMPC/main_parallel.py:10-43
 *   AdaptiveNPMPCSmooth.solve(x0,u_prev,th,R)  RMPC/dev_dual/controller/np_mpc_adaptive_with_linear_regressor.py:212-222
 *   RLS.update(phi, y)                       same file :17-27 (called from RMPC/dev_dual/rob_ctrl.py:340-343)
 *   reference governor + build_ref_traj      rob_ctrl.py:346-351, np_mpc_adaptive...py:201-210
 *   RLMPC._solver_worker NLP solve           LMPC/src/controller/rlmpc2.py:494-519
 *   Policy.mean_net forward + param update   rlmpc2.py:71-80, 742-759, 606-616
 *
 * Each entry point below is the batched form of one of those calls (batch axis B first, B = 1
 * reproduces the reference call).  Plain pointers and sizes only; no torch types.  All arrays are
 * row-major and float64 unless stated.  "dev" pointers are CUDA device pointers, "host" pointers are
 * ordinary host memory.  Every function returns 0 on success or a negative dart_error; none throws,
 * none synchronises the device except the *_host variants.  A handle is bound to one device and must
 * not be used from two threads at once.  There is no CPU implementation behind this ABI: without a
 * CUDA device dart_create fails with DART_ERR_NO_DEVICE.
 */
#ifndef DART_B200_H
#define DART_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DART_PMPC 0
#define DART_RMPC 1
#define DART_LMPC 2

#define DART_STATUS_CONVERGED 0   /* KKT error <= tol                                             */
#define DART_STATUS_MAXITER 1     /* iteration cap hit; last iterate returned (reference: silent)  */
#define DART_STATUS_INFEASIBLE 2  /* x0 violates a stage-0 cap, or the step vanished with violation left   */
#define DART_STATUS_NUMERIC 3     /* NaN/Inf encountered                                           */
#define DART_STATUS_ACCEPTABLE 4  /* acceptable_iter consecutive iterates with error <= acceptable_tol (IPOPT's
                                     "Solved To Acceptable Level"); only when enabled in dart_cfg              */

typedef enum {
    DART_OK = 0,
    DART_ERR_ARG = -1,
    DART_ERR_NO_DEVICE = -2,
    DART_ERR_CUDA = -3,
    DART_ERR_ALLOC = -4,
    DART_ERR_UNSUPPORTED = -5
} dart_error;

/* Problem description = the reference's constructor arguments / parameter dicts.
 * PMPC: mpc_3d.py:12 and PMPC/main.py:59-69.  RMPC: np_mpc_adaptive...py:35-38 and rob_ctrl.py:281-288.
 * LMPC: LMPC/src/run.py:118-126.  Fields of the other methods are ignored. */
typedef struct {
    int32_t method;      /* DART_PMPC | DART_RMPC | DART_LMPC */
    int32_t N;           /* horizon: 15 (PMPC), 20 (RMPC, LMPC) */
    double Ts;           /* model.opt.timestep = 0.002 */
    double g;            /* PMPC: model.opt.gravity[2] = -9.81; RMPC: gz = -9.81; LMPC: unused (literal 9.81) */
    /* bounds */
    double u_lo, u_hi;   /* tilt bounds */
    double du_lo, du_hi; /* RMPC tilt-rate bounds */
    double vmax;         /* RMPC velocity cap */
    double v_eps;        /* RMPC tanh smoothing */
    /* weights */
    double Qp, Qv, R;    /* PMPC (R = tilt weight); RMPC uses Qp, Qv, Ru = R */
    double Rdu;          /* RMPC tilt-rate weight */
    double mu;           /* PMPC viscous friction coefficient */
    double Q[8], Qt[8];  /* LMPC stage / terminal state weights */
    double Rl[4];        /* LMPC [R_a, R_b, R_da, R_db] */
    /* solver options (0 selects the default in brackets) */
    double tol;          /* [1e-8]  KKT tolerance (IPOPT 'tol')            */
    int32_t max_iter;    /* [PMPC 3000->capped 200, RMPC 200, LMPC 200]     */
    double mu_init;      /* [0]     initial barrier parameter; 0 = the barrier strategy's default: 0.1 (IPOPT's mu_init) under the
                          *         monotone schedule, 0.01 (scale of the initial multipliers) with predictor-corrector steps */
    int32_t lanes;       /* [auto]  lanes of a warp cooperating on one sub-problem: 2, 4, 8, 16 or 32 */
    int32_t block_threads; /* [auto] threads per block: a multiple of lcm(32, lanes * axes), axes = 2 for PMPC and LMPC (the two axis problems of an instance share a block), 1 for RMPC; other values -> DART_ERR_ARG */
    /* IPOPT's acceptable-level termination (rlmpc2.py:486-488 sets tol 1e-4, acceptable_tol 1e-3, acceptable_iter 5,
     * max_iter 50 for LMPC; PMPC/RMPC leave IPOPT's 1e-6 / 15, which never triggers before tol on this path).
     * [0 / 0 = off: every solve runs to tol] */
    double acceptable_tol;
    int32_t acceptable_iter;
} dart_cfg;

typedef struct dart_solver* dart_handle;

/* Initial barrier parameter of the solves that follow (0 restores the default 0.1).  A warm-started solve of a
 * closed loop starts close to the previous optimum; beginning the barrier schedule at 1e-4 instead of 0.1 saves the
 * iterations that would only walk mu down (measured: 6.4 -> 4.1 per solve).  The reference's closest knob is
 * IPOPT's mu_init (arm.py:306 sets it for its warm-started worker). */
int dart_set_mu_init(dart_handle h, double mu_init);

/* Barrier-parameter strategy of the solves that follow (IPOPT's `mu_strategy`; the reference leaves IPOPT's default,
 * "monotone", at mpc_3d.py:82 -- the strategy changes the iterate path, not the KKT point a solve converges to).
 * DART_BARRIER_MEHROTRA: Mehrotra predictor-corrector steps with an adaptive barrier parameter wherever the kernel
 * implements them -- PMPC axis problems at the reference horizon (scan sweeps), RMPC and LMPC (tiled sweeps; the corrector
 * re-uses the Riccati factorisation, two extra vector sweeps) -- and the monotone schedule elsewhere;
 * DART_BARRIER_MONOTONE: Fiacco-McCormick schedule (mu0 = mu_init, kappa_mu 0.2, theta_mu 1.5) for every method;
 * DART_BARRIER_AUTO (the default of a new handle): per method, the one measured faster on B200 -- predictor-corrector
 * for PMPC and LMPC; for RMPC predictor-corrector when the call carries no warm plan (warm_w == NULL: 12.95 -> 8.85
 * iterations, 1.43 -> 1.33 ms at 4096 instances) and monotone when it does (the closed loop's warm-started barrier,
 * dart_set_mu_init, needs fewer and cheaper iterations than an adaptive one).
 * (LMPC without a tilt-rate cost, Rl[2] or Rl[3] == 0, always runs the monotone schedule: the corrector of the tiled sweeps
 * recovers its inverse pivots from that coupling.)
 * The environment variables DART_BARRIER_MONOTONE=1 / DART_BARRIER_MEHROTRA=1 change the default of new handles. */
#define DART_BARRIER_MONOTONE 0
#define DART_BARRIER_MEHROTRA 1
#define DART_BARRIER_AUTO 2
int dart_set_barrier_strategy(dart_handle h, int32_t strategy);

/* Dual warm start for closed loops (IPOPT's warm_start_init_point with lam_x0 / lam_g0, which the reference passes
 * for its arm worker at arm.py:420-424 but not for the tray controllers).  `dual` is a DEVICE buffer of
 * [capacity_rows, dart_ndual(h)] doubles that the caller zero-initialises once and then leaves alone: every dart_solve
 * (a) starts instance i from the equality multipliers, slacks and bound multipliers stored in row i -- if that row was
 * written by a solve that ended converged/acceptable and warm_w is given; such a solve starts its barrier at mu_init
 * (use dart_set_mu_init(h, 1e-6)), any other instance never below 1e-4 -- and (b) writes its final slacks and multipliers
 * back to row i.  NULL unregisters.  Rows belong to instances by position, like warm_w. */
int dart_ndual(dart_handle h);
int dart_set_dual_state(dart_handle h, double* dual, int32_t capacity_rows);

/* Fill cfg with the reference's defaults for a method (values cited above). */
int dart_default_cfg(int32_t method, dart_cfg* cfg);

int dart_create(dart_handle* out, const dart_cfg* cfg, int device);
int dart_destroy(dart_handle h);

/* Sizes of the per-instance arrays for this handle. */
int dart_nx(dart_handle h);    /* state size: 6 / 4 / 8                         */
int dart_nref(dart_handle h);  /* target/reference size: 6 / (N+1)*4 / 8          */
int dart_naux(dart_handle h);  /* PMPC 4 [Qp,Qv,R,mu]; RMPC 16 [u_prev,theta_hat]; LMPC 36 [u_prev,pvec] */
int dart_nw(dart_handle h);    /* decision vector size (N+1)*nx + N*2, reference layout w = [vec(X); vec(U)] */

/* One batched NLP solve (replaces the solver call inside PMPC.solve / AdaptiveNPMPCSmooth.solve /
 * _solver_worker).  All pointers are device pointers on the handle's device.
 *   x0      [B, nx]    current state (PMPC: what get_state() returns)
 *   ref     [B, nref]  PMPC target(6) | RMPC Rref_flat((N+1)*4) | LMPC target(8)
 *   aux     [B, naux]  PMPC: NULL -> weights/mu from cfg for every instance; RMPC, LMPC: required
 *   warm_w  [B, nw]    primal warm start in the reference layout, or NULL for the reference's cold start
 *                      (tile(x0) and zeros, mpc_3d.py:123)
 *   w_out   [B, nw]    optimal decision vector (reference's sol['x']); may be NULL
 *   u0_out  [B, 2]     first tilt command U_opt[0]
 *   J_out   [B]        optimal objective (reference's sol['f'])
 *   status  [B] int32  DART_STATUS_*            iters [B] int32 Newton iterations; either may be NULL
 * Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).  The calling thread's current
 * CUDA device must be the handle's device (DART_ERR_ARG otherwise). */
int dart_solve(dart_handle h, int32_t B, const double* x0, const double* ref, const double* aux,
               const double* warm_w, double* w_out, double* u0_out, double* J_out, int32_t* status,
               int32_t* iters, void* stream);

/* Same call with HOST pointers: stages through a pinned buffer, moves inputs to the device, solves, moves results back and
 * synchronises.  This is the call the Python drop-in classes make.  Batches of up to 1 MB without plans (warm_w, w_out NULL)
 * are read from / written to the pinned block by the kernel itself (mapped host memory: the same bytes over the same link,
 * no copy-engine launches); larger ones use device staging and cudaMemcpyAsync.  DART_HOST_STAGED=1 forces the latter. */
int dart_solve_host(dart_handle h, int32_t B, const double* x0, const double* ref, const double* aux,
                    const double* warm_w, double* w_out, double* u0_out, double* J_out,
                    int32_t* status, int32_t* iters);

/* Optional: subsequent dart_solve calls on this handle also write packed result rows [B,4] = [u0x, u0y, J, status]
 * (device pointer, float64, room for capacity_rows rows; NULL switches it off; a solve with B > capacity_rows is
 * refused with DART_ERR_ARG) -- the buffer a multi-GPU caller all-gathers, replacing the
 * (u_cmd, loss, solve_time) tuples of main_parallel.py's control_queue (:43, :201-205). */
int dart_set_result_rows(dart_handle h, double* rows, int32_t capacity_rows);

/* Multi-GPU gather WITHOUT a collective (the instances are independent, so the only exchange of the path is this gather;
 * main_parallel.py:43,201-205 moves the same tuples through mp.Queue): the solve kernel stores every instance's row
 * [u0x, u0y, J, status] directly into the gathered buffers of up to DART_MAX_PEERS GPUs of the node -- peers[p] are
 * peer-mapped device pointers (CUDA IPC / peer access; this GPU's own buffer included), row (row_offset + instance) of each.
 * n_peers = 0 switches it off.  dart_peer_handshake then tells every rank that all rows of a step have landed. */
#define DART_MAX_PEERS 8
int dart_set_result_rows_peers(dart_handle h, double* const* peers, int32_t n_peers, int64_t row_offset);
/* Buffers that other processes of the node can map (CUDA IPC): dart_peer_alloc allocates and zeroes `bytes` on the current
 * device and exports its 64-byte handle; dart_peer_open maps a peer's handle for the CURRENT device (lazy peer access over
 * NVLink), dart_peer_close unmaps it, dart_peer_free releases an own buffer (after every peer has closed it). */
int dart_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle64);
int dart_peer_open(const uint8_t* handle64, void** ptr);
int dart_peer_close(void* ptr);
int dart_peer_free(void* ptr);
/* Enable direct access from `device` to `peer` memory (cudaDeviceEnablePeerAccess; already enabled is not an error). */
int dart_enable_peer_access(int device, int peer);
/* Step hand-shake over peer memory, one launch on `stream`: writes `step` into slot `my_rank` of every peer's flag array
 * (peer_flags[p], int64 [n_peers], peer-mapped) after a system-scope fence, then waits until all n_peers slots of the own array
 * (peer_flags[my_rank]) hold a value >= step.  Work enqueued on the stream after it sees every peer's rows of that step.
 * The wait gives up after about two seconds (a crashed peer must not hang the GPU) and sets *timed_out (device int32). */
int dart_peer_handshake(int64_t* const* peer_flags, int32_t n_peers, int32_t my_rank, int64_t step, int32_t* timed_out,
                        void* stream);

/* Number of kernels launched by this handle since creation (for bench.py's gpu_launches). */
int64_t dart_launch_count(dart_handle h);

/* Launch configuration chosen for the last dart_solve* call. */
int dart_last_launch_config(dart_handle h, int32_t* lanes, int32_t* block_threads, int32_t* grid, int32_t* smem_bytes);

/* tilt -> tray quaternion wxyz, Euler xyz [u1, -u0, 0] (PMPC/main.py:107-116). u [B,2] -> quat [B,4], device. */
int dart_tilt_to_quat(int32_t B, const double* u, double* quat, void* stream);

/* Batched RLS.update (np_mpc_adaptive_with_linear_regressor.py:17-27): E estimators per instance share phi.
 *   theta [B,E,7] in/out   P [B,E,7,7] in/out   phi [B,7]   y [B,E]   lam = forgetting factor.  Device pointers. */
int dart_rls_update(int32_t B, int32_t E, double* theta, double* P, const double* phi, const double* y,
                    double lam, void* stream);

/* One RMPC closed-loop step's pre-solve glue (rob_ctrl.py:335-351) fused into one launch: finite-difference
 * acceleration, regressor of prev_state, the x and y RLS updates, reference governor, build_ref_traj, and the
 * solver's aux rows.  Device pointers:
 *   xk, target [B,4]   prev_state [B,4] in/out (set to xk for the next step, rob_ctrl.py:365; xk != prev_state)
 *   u_prev [B,2] (the last command: may be the u0 buffer of the previous dart_solve)   r_v [B,4] in/out
 *   theta [B,2,7] in/out   P [B,2,7,7] in/out
 *   ref [B,(N+1)*4] out (Rref_flat)   aux [B,16] out ([u_prev, theta_hat] as dart_solve expects for RMPC) */
int dart_rmpc_prologue(int32_t B, int32_t N, double Ts, double v_eps, double lam, double dr_max, double alpha_rg,
                       double step_fraction, const double* xk, double* prev_state, const double* target,
                       const double* u_prev, double* r_v, double* theta, double* P, double* ref, double* aux,
                       void* stream);

/* ---- LMPC parameter-adaptation policy (rlmpc2.py:33-80, 641-668, 742-759, 606-616) ---- */
typedef struct dart_policy* dart_policy_handle;

/* Actor weights as torch stores them (Linear.weight is [out, in], float32, host pointers):
 * W1 [64,520] b1 [64]  W2 [64,64] b2 [64]  W3 [34,64] b3 [34].  Only the reference architecture
 * (obs_dim 520 = 10 x 52 history, hidden 64 x 2, act_dim 34) is supported. */
int dart_policy_create(dart_policy_handle* out, int device, int32_t obs_dim, int32_t hidden, int32_t act_dim,
                       const float* W1, const float* b1, const float* W2, const float* b2, const float* W3,
                       const float* b3);
int dart_policy_destroy(dart_policy_handle h);

/* Policy.mean_net forward: obs [B,520] f32 (device, 16-byte aligned) -> act_mean [B,34] f32 (device).
 * One fused launch.  Default arithmetic DART_POLICY_FP32 matches the reference's FP32 torch forward to <= 2e-5
 * (layer 1: TMA-fed tcgen05 as a 3xTF32 product A_hi W_hi + A_hi W_lo + A_lo W_hi, FP32 accumulate; layers 2, 3: FP32
 * FMAs).  DART_POLICY_TF32 is the single-pass TF32 tensor-core kernel: 1.6x the streaming rate, |error| <= 8e-3. */
#define DART_POLICY_FP32 0
#define DART_POLICY_TF32 1
int dart_policy_set_precision(dart_policy_handle h, int32_t precision);
int dart_policy_forward(dart_policy_handle h, int32_t B, const float* obs, float* act_mean, void* stream);
int64_t dart_policy_launch_count(dart_policy_handle h);

/* Observation build (rlmpc2.py:641-668): base = [state(8), target(8), control(2), cur_k(34)] rounded to f32,
 * Welford update of mean/M2 [B,52] (count = pushes so far including this one), normalise, append to the 10-deep
 * history: obs_out [B,520] = [obs_in[:,52:], normalised base].  obs_in != obs_out.  Device pointers.
 * cur_k has row stride ld_k (34, or 36 when it points into the LMPC aux rows at column 2). */
int dart_policy_obs_push(int32_t B, int32_t count, const double* state, const double* target, const double* control,
                         const double* cur_k, int32_t ld_k, double* mean, double* M2, const float* obs_in,
                         float* obs_out, void* stream);

/* Parameter update (rlmpc2.py:742-759 then write_params_to_shm :606-616): pvec [B, ld_pvec] f64 in/out
 * (the 34 model parameters), action [B,34] f32.  Device pointers. */
int dart_policy_param_update(int32_t B, const float* action, double* pvec, int32_t ld_pvec, double k_max,
                             double max_delta, double min_k, double k_ceiling_margin, double alpha, void* stream);

/* Surrogate closed loop (no MuJoCo): one plant step of the PMPC model (mpc_3d.py:87-104, tilt held over Ts) with
 * per-instance viscous mu [B] and optional unmodelled Coulomb coefficient [B] (NULL = none), plus the episode
 * metrics of PMPC/src/logger.py:155-176 accumulated in place: err [B] = position error of the logged state,
 * conv_time [B] (initialise to -1) = first logged time with err < tol, effort [B] += |u| Ts; nsteps [B] int32
 * (initialise to 0) is the per-instance step counter, kept on the device so that the launch can be replayed from a
 * CUDA graph.  status, iters [B] and counters [2] (uint64) are optional (all or none): the statistics of the solve that
 * produced u are added to counters as dart_pmpc_episode does ([0] += iterations, [1] += solves not converged).
 * Device pointers. */
int dart_pmpc_plant_step(int32_t B, double Ts, double g, const double* mu, const double* coulomb, const double* u,
                         const double* target, double* state, int32_t* nsteps, double tol, double* conv_time,
                         double* effort, double* err, const int32_t* status, const int32_t* iters, uint64_t* counters,
                         void* stream);

/* Low-level arm controller QP: replaces the per-cycle ca.nlpsol('solver','ipopt',...) construction and call of
 * ARMCONTROL.solver_worker, PMPC/src/controller/arm.py:337-457, for B arms at once:
 *     min 0.5 x'Hx + g'x   s.t.  lo <= Cx <= hi,     x = joint accelerations qdd (7),
 *     C = [0.5 dt^2 I; dt I; M] (21 rows: joint-position, joint-velocity and torque limits, arm.py:399-405).
 * H [B,7,7] (symmetric positive definite), g [B,7], C [B,21,7], lo/hi [B,21], x0 [B,7] (optional primal start, the
 * reference's prev_qdd) -> x [B,7], obj [B] = 0.5 x'Hx + g'x, status [B] (DART_STATUS_*), iters [B]; all device pointers,
 * row-major f64.  tol <= 0 selects 1e-8, max_iter <= 0 selects 100.  The caller forms H, g, lo, hi from the MuJoCo
 * quantities (dart_b200.arm.build_qp mirrors arm.py:337-405). */
int dart_arm_qp_solve(int32_t B, const double* H, const double* g, const double* C, const double* lo, const double* hi,
                      const double* x0, double* x, double* obj, int32_t* status, int32_t* iters, double tol,
                      int32_t max_iter, void* stream);
/* Forms the QP data above on the device from what ARMCONTROL.compute_dynamics returns (arm.py:186-200) and the controller
 * parameters (arm.py:495-519) -- the numpy block arm.py:337-405: pinv(M), inv/pinv of the task-space inertia Mx_inv, its
 * matrix square root, damping D, impedance force F, null-space target beta.  Parameters are HOST pointers (row-major:
 * Wimp, K [6,6]; Wpos, Wsmooth, K_null [7,7]; limits_lo/hi [21] = [Qmin|Qdotmin|taumin], [Qmax|Qdotmax|taumax]); per-arm
 * inputs and outputs are DEVICE pointers: q, qd, qdd_prev, h [B,7]; mocap_pos, ee_pos, rotvec [B,3]; jac, jacDot [B,6,7];
 * M [B,7,7]; Mx_inv [B,6,6] -> H, g, c0 [B] (cost = 0.5 x'Hx + g'x + c0 = the reference's loss), C, lo, hi. */
int dart_arm_qp_build(int32_t B, const double* Wimp, const double* Wpos, const double* Wsmooth, const double* K,
                      const double* K_null, const double* limits_lo, const double* limits_hi, double dt,
                      const double* q, const double* qd, const double* qdd_prev, const double* mocap_pos,
                      const double* ee_pos, const double* rotvec, const double* jac, const double* jacDot,
                      const double* M, const double* h, const double* Mx_inv, double* H, double* g, double* c0,
                      double* C, double* lo, double* hi, void* stream);
int64_t dart_arm_qp_launch_count(void);

/* T closed-loop steps of B PMPC instances in ONE launch: each step = dart_solve (cold start, as the reference) followed by
 * dart_pmpc_plant_step, with the state fed back on the device.  Same arithmetic as calling the two T times (bit-identical
 * states and metrics); removes the per-step launch gaps that bound the small-batch closed loop.  `state` [B,6] is read and
 * advanced in place; u0/J/status/iters hold the last step's solve; counters[0] += Newton iterations of all solves,
 * counters[1] += solves that did not end converged (uint64, device).  PMPC handles with the reference horizon only. */
int dart_pmpc_episode(dart_handle h, int32_t B, int32_t T, double* state, const double* target, const double* aux,
                      const double* mu_plant, const double* coulomb, double tol, int32_t* nsteps, double* conv_time,
                      double* effort, double* err, double* u0, double* J, int32_t* status, int32_t* iters,
                      uint64_t* counters, void* stream);

/* RMPC surrogate plant (no reference counterpart -- the reference steps MuJoCo, rob_ctrl.py:365):
 * v' = gz sin(u) - mu |g| tanh(v / 0.01) - c v per axis, four explicit Euler sub-steps of Ts / 4.  mu, c [B]; u [B,2];
 * state [B,4] -> state_out [B,4] (may alias state). */
int dart_rmpc_plant_step(int32_t B, double Ts, double gz, const double* mu, const double* c, const double* u,
                         const double* state, double* state_out, void* stream);

/* LMPC surrogate plant: one RK4 step (tilt u [B,2] held over Ts) of the 8-state model of
 * rlmpc2.py:260-436 with per-instance TRUE parameters: true_aux [B,36] has the layout of the LMPC solver's aux rows
 * ([u_prev(2), pvec(34)]; the first two columns are ignored).  state [B,8] -> state_out [B,8] (may alias).  With
 * true_aux equal to the controller's aux the result is the controller's own one-step prediction, bit for bit. */
/* End of one LMPC closed-loop step in ONE launch.  Always: u_prev [B,2] <- u0 and aux[:,0:2] <- u0 (aux [B,36] are the LMPC
 * solver's rows; rlmpc2.py:1019-1021 `last_control`, views['control']).  With plan_U != NULL also the facade's "no fresh
 * solution" branch (RLMPC.solve, rlmpc2.py:1013-1018) in deterministic form: an instance whose solve did not end converged /
 * acceptable (status [B]), or whose fresh [B] flag (uint8, optional) is 0, gets the next entry of its last good plan as u0
 * (plan_U [B,N,2], plan_pos [B] int64, have_plan [B] uint8; last_control while there is no plan yet), keeps that plan as the
 * warm start (w_next [B,nw] <- w_prev) and adds 1 to n_fallback (uint64); otherwise the new plan is taken over. */
int dart_lmpc_post_step(int32_t B, int32_t N, const int32_t* status, const uint8_t* fresh, const double* w_prev,
                        double* w_next, double* u0, double* u_prev, double* aux, double* plan_U, int64_t* plan_pos,
                        uint8_t* have_plan, uint64_t* n_fallback, void* stream);

int dart_lmpc_plant_step(int32_t B, double Ts, const double* true_aux, const double* u, const double* state,
                         double* state_out, void* stream);

/* ---- PPO training of the LMPC policy (RLMPC._rl_worker, LMPC/src/controller/rlmpc2.py:536-935) ----
 * The reference trains one instance's policy with torch autograd on whatever device torch finds; these entry points run the
 * same arithmetic (FP32 networks, Adam, clipped surrogate) for B instances that share one policy, without torch.
 * Flat parameter vector, float32, dart_ppo_nparams() = 77 317 entries, torch layout [out,in] per matrix:
 *   [W1 (128x520: rows 0-63 mean_net.0.weight, rows 64-127 value_net.0.weight) | b1 (128) |
 *    mean_net.2.weight (64x64) | .bias (64) | value_net.2.weight (64x64) | .bias (64) |
 *    mean_net.4.weight (34x64) | .bias (34) | value_net.4.weight (1x64) | .bias (1) | log_std (34)]            */
typedef struct dart_ppo* dart_ppo_handle;

typedef struct dart_ppo_cfg {
    double lr, weight_decay, beta1, beta2, adam_eps;      /* optim.Adam(lr, weight_decay=1e-5)        rlmpc2.py:561     */
    double clip_eps, vf_coef, ent_coef, max_grad_norm;    /* loss and clip_grad_norm_(0.5)            rlmpc2.py:806-816 */
    double log_std_min, log_std_max;                      /* Policy.forward clamp                     rlmpc2.py:60-61,76 */
} dart_ppo_cfg;

typedef struct dart_ppo_reward_cfg {                      /* rlmpc2.py:598-601, 701-735                                  */
    double max_delta, action_scale, max_per_dim_rms;      /* delta_z = a * max_delta * action_scale, damped above the rms */
    double sigma_pos, sigma_vel, w_pos, w_vel, w_change, w_d_ctrl;
    double success_tol, success_bonus, oob_penalty, no_contact_penalty;
    double tray_limit[2];
    int32_t max_episode_steps;
    double time_penalty_inc;
} dart_ppo_reward_cfg;

int dart_ppo_default_cfg(dart_ppo_cfg* cfg);
int dart_ppo_default_reward_cfg(dart_ppo_reward_cfg* cfg);
int dart_ppo_nparams(void);
/* capacity = largest B of dart_ppo_act and largest minibatch of dart_ppo_update.  params_host: the flat vector above. */
int dart_ppo_create(dart_ppo_handle* out, int device, int32_t obs_dim, int32_t hidden, int32_t act_dim, int32_t capacity,
                    const float* params_host, const dart_ppo_cfg* cfg);
int dart_ppo_destroy(dart_ppo_handle h);
/* Checkpointing (torch.save({"model", "optimizer", ...}), rlmpc2.py:920-922): parameters, Adam moments (host pointers,
 * each nullable) and the optimiser step count.  Both calls synchronise the device. */
int dart_ppo_get_state(dart_ppo_handle h, float* params_host, float* m_host, float* v_host, int64_t* step);
int dart_ppo_set_state(dart_ppo_handle h, const float* params_host, const float* m_host, const float* v_host, int64_t step);
int dart_ppo_get_grad(dart_ppo_handle h, float* grad_host);          /* last dart_ppo_update's gradient, before clipping */
const float* dart_ppo_params_dev(dart_ppo_handle h);                 /* device pointer of the live parameters           */
int64_t dart_ppo_launch_count(dart_ppo_handle h);
/* Rollout-time policy call (rlmpc2.py:670-699): mean, std, value = policy(obs); action = mean + std * eps (eps [B,34]
 * standard normal draws supplied by the caller, NULL = the mean action), logp = sum_j log N(action_j).  obs [B,520],
 * action/mean [B,34] (mean nullable), logp/value [B]; float32 device pointers. */
int dart_ppo_act(dart_ppo_handle h, int32_t B, const float* obs, const float* eps, float* action, float* logp,
                 float* value, float* mean, void* stream);
/* Reward and termination of one control step for B instances (rlmpc2.py:701-735).  state/target [B,8], control [B,2],
 * prev_cmd [B,2] (in/out), in_contact [B] (nullable = in contact) f64; action [B,34] f32 (the raw action);
 * episode_step [B] int32 and time_penalty [B] f64 are advanced in place and reset where done; reward/done [B] f32. */
int dart_ppo_reward(int32_t B, const dart_ppo_reward_cfg* cfg, const double* state, const double* target,
                    const double* control, double* prev_cmd, const float* action, const double* in_contact,
                    int32_t* episode_step, double* time_penalty, float* reward, float* done, void* stream);
/* GAE along each instance's rollout (compute_gae, rlmpc2.py:589-596) and returns = adv + values (:784).  Arrays [T,B] f32. */
int dart_ppo_gae(int32_t B, int32_t T, const float* rewards, const float* values, const float* dones,
                 const float* last_value, double gamma, double lam, float* adv, float* ret, void* stream);
/* In place x = (x - mean) / (std + 1e-8) over n values; ddof 0 = numpy std (returns, :785), 1 = torch std (advantages, :792). */
int dart_ppo_normalize(int64_t n, float* x, int32_t ddof, void* stream);
/* One minibatch step (rlmpc2.py:797-817): forward of both networks, clipped-surrogate + vf_coef * MSE - ent_coef * entropy,
 * backward, clip_grad_norm_, Adam.  idx [M] int64 (nullable) gathers the minibatch rows from obs [*,520] (16-byte
 * aligned), act [*,34], old_logp, adv, ret [*].  apply = 0 computes the gradient only.  stats [4] (nullable, device) =
 * policy loss, value loss, entropy, gradient norm before clipping.  Results are bitwise repeatable.  The optimiser step
 * count lives on the device and no argument changes from step to step, so the launches may be captured in a CUDA graph. */
int dart_ppo_update(dart_ppo_handle h, int32_t M, const int64_t* idx, const float* obs, const float* act,
                    const float* old_logp, const float* adv, const float* ret, int32_t apply, float* stats, void* stream);

/* Data-parallel training (one process per GPU, shared policy): after dart_ppo_update(..., apply = 0) each rank exports its
 * minibatch gradient [dart_ppo_nparams()] to a device buffer the host side all-reduces (NCCL over NVLink: the one exchange step of
 * this path, 309 kB per step), then applies scale * sum (scale = 1 / world size) with the usual global-norm clip + Adam.  Every
 * rank applies the same bits, so the replicas stay identical without a parameter broadcast. */
int dart_ppo_export_grad(dart_ppo_handle h, float* dst_dev, void* stream);
int dart_ppo_apply_grad(dart_ppo_handle h, const float* src_dev, double scale, float* stats, void* stream);

/* Measured FP64 FMA-pipe peak of the device in TFLOP/s (DFMA microbenchmark, CUDA-event timed): the roofline
 * denominator bench.py reports the solver kernels against. */
int dart_measure_fp64_tflops(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* DART_B200_H */
