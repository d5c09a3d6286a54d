"""ctypes binding of the C ABI in include/dart_b200.h (lib/libdart_b200.so, built by __graft_entry__.build()).

There is no CPU implementation behind this module: if the shared library is missing, or no CUDA
device is present, every entry point raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DART_B200_LIB") or os.path.join(HERE, "lib", "libdart_b200.so")   # override: A/B builds

DART_PMPC, DART_RMPC, DART_LMPC = 0, 1, 2
STATUS_CONVERGED, STATUS_MAXITER, STATUS_INFEASIBLE, STATUS_NUMERIC, STATUS_ACCEPTABLE = 0, 1, 2, 3, 4
ERRORS = {0: "ok", -1: "bad argument", -2: "no CUDA device", -3: "CUDA error", -4: "allocation failed", -5: "unsupported configuration"}


class DartCfg(C.Structure):
    """Mirror of ``dart_cfg`` (include/dart_b200.h)."""
    _fields_ = [
        ("method", C.c_int32), ("N", C.c_int32), ("Ts", C.c_double), ("g", C.c_double),
        ("u_lo", C.c_double), ("u_hi", C.c_double), ("du_lo", C.c_double), ("du_hi", C.c_double),
        ("vmax", C.c_double), ("v_eps", C.c_double),
        ("Qp", C.c_double), ("Qv", C.c_double), ("R", C.c_double), ("Rdu", C.c_double), ("mu", C.c_double),
        ("Q", C.c_double * 8), ("Qt", C.c_double * 8), ("Rl", C.c_double * 4),
        ("tol", C.c_double), ("max_iter", C.c_int32), ("mu_init", C.c_double),
        ("lanes", C.c_int32), ("block_threads", C.c_int32),
        ("acceptable_tol", C.c_double), ("acceptable_iter", C.c_int32),
    ]


class DartError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libdart_b200.so once; raise loudly when it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DartError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). dart_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
    L.dart_default_cfg.argtypes = [C.c_int32, C.POINTER(DartCfg)]
    L.dart_create.argtypes = [C.POINTER(vp), C.POINTER(DartCfg), C.c_int]
    L.dart_destroy.argtypes = [vp]
    for f in ("dart_nx", "dart_nref", "dart_naux", "dart_nw"):
        getattr(L, f).argtypes = [vp]
    L.dart_solve.argtypes = [vp, C.c_int32] + [vp] * 9 + [vp]
    L.dart_solve_host.argtypes = [vp, C.c_int32] + [vp] * 9
    if hasattr(L, "dart_set_result_rows"):          # absent only in older A/B builds loaded through DART_B200_LIB
        L.dart_set_result_rows.argtypes = [vp, vp, C.c_int32]
    if hasattr(L, "dart_set_mu_init"):
        L.dart_set_mu_init.argtypes = [vp, C.c_double]
    L.dart_set_barrier_strategy.argtypes = [vp, C.c_int32]
    if hasattr(L, "dart_set_dual_state"):
        L.dart_set_dual_state.argtypes = [vp, vp, C.c_int32]
        L.dart_ndual.argtypes = [vp]
    if hasattr(L, "dart_set_result_rows_peers"):
        L.dart_set_result_rows_peers.argtypes = [vp, C.POINTER(vp), C.c_int32, C.c_int64]
        L.dart_enable_peer_access.argtypes = [C.c_int, C.c_int]
        L.dart_peer_alloc.argtypes = [C.c_int64, C.POINTER(vp), C.c_char_p]
        L.dart_peer_open.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.dart_peer_close.argtypes = [vp]
        L.dart_peer_free.argtypes = [vp]
        L.dart_peer_handshake.argtypes = [C.POINTER(vp), C.c_int32, C.c_int32, C.c_int64, vp, vp]
    L.dart_launch_count.argtypes = [vp]
    L.dart_launch_count.restype = C.c_int64
    L.dart_last_launch_config.argtypes = [vp, ip, ip, ip, ip]
    L.dart_tilt_to_quat.argtypes = [C.c_int32, vp, vp, vp]
    L.dart_rls_update.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_double, vp]
    L.dart_rmpc_prologue.argtypes = [C.c_int32, C.c_int32] + [C.c_double] * 6 + [vp] * 9 + [vp]
    fp = C.POINTER(C.c_float)
    L.dart_policy_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int32, C.c_int32, C.c_int32] + [vp] * 6
    L.dart_policy_destroy.argtypes = [vp]
    L.dart_policy_forward.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.dart_policy_set_precision.argtypes = [vp, C.c_int32]
    L.dart_policy_launch_count.argtypes = [vp]
    L.dart_policy_launch_count.restype = C.c_int64
    L.dart_policy_obs_push.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp]
    L.dart_policy_param_update.argtypes = [C.c_int32, vp, vp, C.c_int32] + [C.c_double] * 5 + [vp]
    if hasattr(L, "dart_pmpc_plant_step"):
        L.dart_pmpc_plant_step.argtypes = [C.c_int32, C.c_double, C.c_double] + [vp] * 6 + [C.c_double] + [vp] * 6 + [vp]
    L.dart_measure_fp64_tflops.argtypes = [C.c_int, dp]
    if hasattr(L, "dart_pmpc_episode"):
        L.dart_pmpc_episode.argtypes = [vp, C.c_int32, C.c_int32] + [vp] * 5 + [C.c_double] + [vp] * 9 + [vp]
    if hasattr(L, "dart_arm_qp_solve"):
        L.dart_arm_qp_solve.argtypes = [C.c_int32] + [vp] * 10 + [C.c_double, C.c_int32, vp]
        L.dart_arm_qp_launch_count.restype = C.c_int64
        L.dart_arm_qp_build.argtypes = [C.c_int32] + [vp] * 7 + [C.c_double] + [vp] * 17 + [vp]
    if hasattr(L, "dart_rmpc_plant_step"):
        L.dart_rmpc_plant_step.argtypes = [C.c_int32, C.c_double, C.c_double, vp, vp, vp, vp, vp, vp]
    if hasattr(L, "dart_lmpc_plant_step"):
        L.dart_lmpc_plant_step.argtypes = [C.c_int32, C.c_double, vp, vp, vp, vp, vp]
        L.dart_lmpc_post_step.argtypes = [C.c_int32, C.c_int32] + [vp] * 11 + [vp]
    if hasattr(L, "dart_ppo_create"):
        L.dart_ppo_default_cfg.argtypes = [vp]
        L.dart_ppo_default_reward_cfg.argtypes = [vp]
        L.dart_ppo_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp]
        L.dart_ppo_destroy.argtypes = [vp]
        L.dart_ppo_get_state.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int64)]
        L.dart_ppo_set_state.argtypes = [vp, vp, vp, vp, C.c_int64]
        L.dart_ppo_get_grad.argtypes = [vp, vp]
        L.dart_ppo_params_dev.argtypes = [vp]
        L.dart_ppo_params_dev.restype = vp
        L.dart_ppo_launch_count.argtypes = [vp]
        L.dart_ppo_launch_count.restype = C.c_int64
        L.dart_ppo_act.argtypes = [vp, C.c_int32] + [vp] * 6 + [vp]
        L.dart_ppo_reward.argtypes = [C.c_int32] + [vp] * 11 + [vp]
        L.dart_ppo_gae.argtypes = [C.c_int32, C.c_int32] + [vp] * 4 + [C.c_double, C.c_double, vp, vp, vp]
        L.dart_ppo_normalize.argtypes = [C.c_int64, vp, C.c_int32, vp]
        L.dart_ppo_update.argtypes = [vp, C.c_int32] + [vp] * 6 + [C.c_int32, vp, vp]
        if hasattr(L, "dart_ppo_apply_grad"):
            L.dart_ppo_export_grad.argtypes = [vp, vp, vp]
            L.dart_ppo_apply_grad.argtypes = [vp, vp, C.c_double, vp, vp]
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise DartError(f"{what} failed: {ERRORS.get(rc, rc)} ({rc})")
