"""Problem configuration: the reference's hard-coded parameter dicts as a YAML schema.

The reference README advertises a ``config.yaml`` per method but ships none; every value is a literal
in PMPC/main.py:59-69, PMPC/main_parallel.py:107-122, RMPC/dev_dual/rob_ctrl.py:281-288 and
LMPC/src/run.py:118-151.  ``config.yaml`` at the repo root carries exactly those keys; this module
turns a method section into the C ABI's ``dart_cfg``.
"""
import os

from ._lib import DART_LMPC, DART_PMPC, DART_RMPC, DartCfg

METHODS = {"pmpc": DART_PMPC, "rmpc": DART_RMPC, "lmpc": DART_LMPC}

_SOLVER_KEYS = ("tol", "max_iter", "mu_init", "lanes", "block_threads", "acceptable_tol", "acceptable_iter")

# The IPOPT options the reference passes for LMPC (rlmpc2.py:484-488; max_cpu_time has no counterpart here).  Not the
# default: by default every method is solved to tol = 1e-8 so results do not depend on where an early exit lands.
LMPC_REFERENCE_SOLVER_OPTIONS = dict(tol=1e-4, acceptable_tol=1e-3, acceptable_iter=5, max_iter=50)


def pmpc_cfg(Ts=0.002, nx=6, nu=2, N=15, Qp=400.0, Qv=2.0, R=0.2, mu=0.1, u_bounds=(-0.6, 0.6), g=-9.81, **solver):
    """PMPC/main.py:59-69 defaults (the class defaults of mpc_3d.py:12 are N=20, Qp=100, Qv=0, R=0.1, mu=0.4, +-0.5)."""
    if nx != 6 or nu != 2:
        raise ValueError("PMPC is defined for nx=6, nu=2 (mpc_3d.py:87-97)")
    c = DartCfg()
    c.method, c.N, c.Ts, c.g = DART_PMPC, int(N), float(Ts), float(g)
    c.u_lo, c.u_hi = float(u_bounds[0]), float(u_bounds[1])
    c.Qp, c.Qv, c.R, c.mu = float(Qp), float(Qv), float(R), float(mu)
    return _solver(c, solver)


def rmpc_cfg(Ts=0.002, nx=4, nu=2, N=20, Qp=80.0, Qv=2.0, Ru=0.02, Rdu=1.0, u_bounds=(-0.6, 0.6),
             du_bounds=(-0.06, 0.06), vmax=0.2, v_eps=0.1, gz=-9.81, **solver):
    """RMPC/dev_dual/rob_ctrl.py:281-284."""
    if nx != 4 or nu != 2:
        raise ValueError("RMPC is defined for nx=4, nu=2")
    c = DartCfg()
    c.method, c.N, c.Ts, c.g = DART_RMPC, int(N), float(Ts), float(gz)
    c.u_lo, c.u_hi = float(u_bounds[0]), float(u_bounds[1])
    c.du_lo, c.du_hi = float(du_bounds[0]), float(du_bounds[1])
    c.vmax, c.v_eps = float(vmax), float(v_eps)
    c.Qp, c.Qv, c.R, c.Rdu = float(Qp), float(Qv), float(Ru), float(Rdu)
    return _solver(c, solver)


def lmpc_cfg(Ts=0.002, nx=8, nu=2, N=20, Q=(200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0),
             Qt=(200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0), R=(0.1, 0.1, 1.0, 1.0), u_bounds=(-0.4, 0.4),
             g=9.81, **solver):
    """LMPC/src/run.py:118-126."""
    if nx != 8 or nu != 2:
        raise ValueError("LMPC is defined for nx=8, nu=2")
    c = DartCfg()
    c.method, c.N, c.Ts, c.g = DART_LMPC, int(N), float(Ts), float(g)
    c.u_lo, c.u_hi = float(u_bounds[0]), float(u_bounds[1])
    for i in range(8):
        c.Q[i], c.Qt[i] = float(Q[i]), float(Qt[i])
    for i in range(4):
        c.Rl[i] = float(R[i])
    return _solver(c, solver)


def _solver(c, solver):
    c.tol = float(solver.pop("tol", 1e-8))
    c.max_iter = int(solver.pop("max_iter", 200))
    c.mu_init = float(solver.pop("mu_init", 0.0))      # 0: the barrier strategy's default (0.1 monotone, 0.01 predictor-corrector)
    c.lanes = int(solver.pop("lanes", 0))
    c.block_threads = int(solver.pop("block_threads", 0))
    c.acceptable_tol = float(solver.pop("acceptable_tol", 0.0))
    c.acceptable_iter = int(solver.pop("acceptable_iter", 0))
    # keys of the reference dicts that do not enter the NLP are accepted and ignored
    return c


BUILDERS = {"pmpc": pmpc_cfg, "rmpc": rmpc_cfg, "lmpc": lmpc_cfg}


def load_config(path=None):
    """Read config.yaml -> {method: dict}.  Default: the file at the repo root."""
    import yaml
    if path is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config.yaml")
    with open(path) as fh:
        return yaml.safe_load(fh)


def cfg_from_yaml(method, path=None, **override):
    """Build the dart_cfg of ``method`` ('pmpc' | 'rmpc' | 'lmpc') from config.yaml (+ overrides)."""
    doc = load_config(path)
    sec = dict(doc.get(method, {}))
    nlp = dict(sec.get("nlp", {}))
    nlp.update(sec.get("solver", {}))
    nlp.update(override)
    import inspect
    allowed = set(inspect.signature(BUILDERS[method]).parameters) | set(_SOLVER_KEYS)
    return BUILDERS[method](**{k: v for k, v in nlp.items() if k in allowed})
