"""Load the reference's hot-path source files by path, unmodified, with stand-ins for what is not installed.

TEST INFRASTRUCTURE ONLY, and usable only where ``/root/reference`` exists (this container; not the GPU box).  The
fixtures it produces (``tests/golden/make_ref_golden.py`` -> ``tests/golden/ref_*.npz``) are what travels.

Stand-ins: ``casadi`` -> ``oracle.refshim.casadi`` (graph-recording, see there); ``mujoco`` -> duck-typed model/data
(the controllers read ``model.opt.gravity[2]``, ``model.opt.timestep``, ``data.body(name).xpos/.cvel/.xmat/.cacc``,
``mujoco.mj_name2id``); ``icecream``, ``cvxpy``, ``matplotlib`` -> empty modules with the imported names.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("DART_REFERENCE_ROOT", "/root/reference")

PMPC_FILE = "PMPC/src/controller/mpc_3d.py"
RMPC_FILE = "RMPC/dev_dual/controller/np_mpc_adaptive_with_linear_regressor.py"
LMPC_FILE = "LMPC/src/controller/rlmpc2.py"


def available():
    return os.path.isfile(os.path.join(REF_ROOT, PMPC_FILE))


# ---- duck-typed MuJoCo ------------------------------------------------------------------------------------------------
class _Body:
    def __init__(self, bid=1):
        self.id = bid
        self.xpos = np.zeros(3)
        self.cvel = np.zeros(6)
        self.cacc = np.zeros(6)
        self.xmat = np.eye(3).reshape(-1)


class FakeData:
    """``data.body(name)`` -> object with xpos (3), cvel (6: angular, linear), cacc, xmat."""

    def __init__(self):
        self._bodies = {}
        self.ncon = 0
        self.contact = []

    def body(self, name):
        return self._bodies.setdefault(name, _Body(len(self._bodies) + 1))


class FakeModel:
    def __init__(self, gravity_z=-9.81, timestep=0.002):
        self.opt = types.SimpleNamespace(gravity=np.array([0.0, 0.0, gravity_z]), timestep=timestep)
        self._bodies = {}

    def body(self, name):
        return self._bodies.setdefault(name, _Body(len(self._bodies) + 1))


def _stub_modules():
    from . import casadi as shim
    mods = {"casadi": shim}
    mj = types.ModuleType("mujoco")
    mj.MjModel = FakeModel
    mj.MjData = FakeData
    mj.mjtObj = types.SimpleNamespace(mjOBJ_BODY=1)
    mj.mj_name2id = lambda model, typ, name: 1
    mj.mj_step = lambda m, d: None
    mjv = types.ModuleType("mujoco.viewer")
    mjv.launch_passive = lambda *a, **k: None
    mj.viewer = mjv
    mods["mujoco"], mods["mujoco.viewer"] = mj, mjv
    ic = types.ModuleType("icecream")
    ic.ic = lambda *a, **k: None
    mods["icecream"] = ic
    cv = types.ModuleType("cvxpy")
    cv.pos = lambda x: x
    mods["cvxpy"] = cv
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        mods["matplotlib"], mods["matplotlib.pyplot"] = mpl, plt
    return mods


_loaded = {}


def load(relpath):
    """Execute a reference source file as a module (cached).  The stand-in modules are visible only during the import."""
    if relpath in _loaded:
        return _loaded[relpath]
    path = os.path.join(REF_ROOT, relpath)
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    stubs = _stub_modules()
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        name = "_dart_ref_" + os.path.splitext(os.path.basename(relpath))[0]
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _loaded[relpath] = mod
    return mod


def nested_function(relpath, outer_path, name, env=None):
    """Compile ONE nested function of a reference file from its own source text (e.g. ``compute_gae`` inside
    ``RLMPC._rl_worker``, rlmpc2.py:589-596), with ``env`` as its globals (closure variables are supplied there).

    outer_path: names of the enclosing class / functions, e.g. ("RLMPC", "_rl_worker")."""
    path = os.path.join(REF_ROOT, relpath)
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    node = tree
    for nm in tuple(outer_path) + (name,):
        found = None
        for child in ast.walk(node):
            if child is not node and isinstance(child, (ast.FunctionDef, ast.ClassDef)) and child.name == nm:
                found = child
                break
        if found is None:
            raise LookupError(f"{nm} not found under {'.'.join(outer_path)} in {relpath}")
        node = found
    modast = ast.Module(body=[node], type_ignores=[])
    code = compile(modast, path, "exec")
    g = {"np": np, "__builtins__": __builtins__}
    g.update(env or {})
    exec(code, g)
    return g[name], (node.lineno, node.end_lineno)
