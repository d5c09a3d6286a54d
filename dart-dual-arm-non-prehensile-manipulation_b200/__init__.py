"""dart_b200 -- B200-native batched nonlinear MPC for DART's tray-tilt high-level controller.

The directory is named ``dart-dual-arm-non-prehensile-manipulation_b200``; import it as ``dart_b200``
through the shim at the repository root (``dart_b200.py``).
"""
from . import _lib, config, workloads                     # noqa: F401
from ._lib import (DART_LMPC, DART_PMPC, DART_RMPC, STATUS_CONVERGED, STATUS_INFEASIBLE,  # noqa: F401
                   STATUS_MAXITER, STATUS_NUMERIC, STATUS_ACCEPTABLE, DartCfg, DartError)
from .config import cfg_from_yaml, lmpc_cfg, load_config, pmpc_cfg, rmpc_cfg   # noqa: F401
from .engine import NMPCEngine, measure_fp64_tflops, tilt_to_quat_device        # noqa: F401
from .pmpc import PMPC, GravityModel, StateHolder, mpc_worker   # noqa: F401
from .rmpc import RLS, AdaptiveNPMPCSmooth, RMPCBatch, rls_update_device, rmpc_plant_step_device   # noqa: F401
from .lmpc import RLMPC, LMPCBatch, PolicyMLP, lmpc_plant_step, init_policy_weights, load_checkpoint_weights, load_checkpoint   # noqa: F401
from .parallel import ShardedSolver, shard_bounds   # noqa: F401
from .episodes import PMPCEpisodes   # noqa: F401
from .arm import ARMCONTROL, ArmQPBatch   # noqa: F401
from . import arm   # noqa: F401
from .ppo import LMPCTrainer, PPOTrainer, init_policy_state, pack_params, unpack_params   # noqa: F401
from . import ppo   # noqa: F401
