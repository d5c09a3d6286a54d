"""Drop-in for the reference's physics-based controller ``PMPC`` and its worker loop.

Mirrors PMPC/src/controller/mpc_3d.py:11-138 (constructor signature, public attributes,
``solve(target) -> (u0 (2,), loss (1,))``, ``get_state()``) and the queue protocol of
PMPC/main_parallel.py:10-43, with the CasADi/IPOPT solve replaced by the CUDA engine.
``model`` is only read for ``model.opt.gravity[2]`` and ``data`` for
``data.body(name).xpos / .cvel`` -- duck-typed objects suffice, MuJoCo is not required.
A batch axis is added through ``solve_batch``.
"""
import time

import numpy as np

from .config import pmpc_cfg
from .engine import NMPCEngine


class PMPC:
    def __init__(self, model, data, Ts, nx=6, nu=2, N=20, Qp=100, Qv=0, R=0.1, mu=0.4, u_bounds=(-0.5, 0.5),
                 device=0, **solver):
        self.model = model
        self.data = data
        self.Ts = Ts
        self.nx = nx
        self.nu = nu
        self.N = N
        self.Qp = Qp
        self.Qv = Qv
        self.R = R
        self.mu = mu
        self.g = float(model.opt.gravity[2])          # mpc_3d.py:23 (negative)
        self.h_cube = 0.1
        self.u_bounds = u_bounds
        self.target_body = "cube"
        self._engine = NMPCEngine(pmpc_cfg(Ts=Ts, nx=nx, nu=nu, N=N, Qp=Qp, Qv=Qv, R=R, mu=mu, u_bounds=u_bounds,
                                           g=self.g, **solver), device=device)
        self.w0 = np.zeros(nx * (N + 1) + nu * N)
        self.status = None
        self.iters = None

    def get_state(self):
        """[px, vx, py, vy, pz, vz] of ``target_body`` (mpc_3d.py:106-113)."""
        pos = self.data.body(self.target_body).xpos
        vel = self.data.body(self.target_body).cvel[3:6]
        return np.array([pos[0], vel[0], pos[1], vel[1], pos[2], vel[2]])

    def solve(self, target):
        """One NLP solve from the reference's cold start; returns (U_opt[0], loss) as mpc_3d.py:133-138."""
        state = self.get_state()
        out = self._engine.solve(state[None, :], np.asarray(target, dtype=np.float64)[None, :])
        self.w0 = out["w"][0]
        self.status = int(out["status"][0])
        self.iters = int(out["iters"][0])
        return out["u0"][0].copy(), out["J"].copy()

    def solve_batch(self, states, targets, params=None, want_w=False):
        """B independent solves: states/targets [B,6]; params [B,4] = per-instance (Qp, Qv, R, mu) or None."""
        return self._engine.solve(states, targets, aux=params, want_w=want_w)

    @property
    def engine(self):
        return self._engine


class _Body:
    def __init__(self):
        self.xpos = np.zeros(3)
        self.cvel = np.zeros(6)


class StateHolder:
    """Minimal stand-in for MuJoCo's ``data``: ``data.body(name).xpos / .cvel`` (what PMPC reads)."""

    def __init__(self):
        self._bodies = {}

    def body(self, name):
        return self._bodies.setdefault(name, _Body())


class GravityModel:
    """Minimal stand-in for MuJoCo's ``model``: ``model.opt.gravity`` and ``model.opt.timestep``."""

    class _Opt:
        def __init__(self, g, ts):
            self.gravity = np.array([0.0, 0.0, g])
            self.timestep = ts

    def __init__(self, g=-9.81, timestep=0.002):
        self.opt = GravityModel._Opt(g, timestep)


def mpc_worker(model_path, target_body, params, state_queue, control_queue, device=0):
    """Queue service loop of PMPC/main_parallel.py:10-43: items ``(state, target)`` or ``"STOP"``;
    replies ``(u_cmd, loss, solve_time)``.  ``model_path`` may be a MuJoCo XML path (needs mujoco) or a
    ``(model, data)`` pair of duck-typed objects."""
    if isinstance(model_path, (tuple, list)):
        model, data = model_path
    else:
        import mujoco
        model = mujoco.MjModel.from_xml_path(model_path)
        data = mujoco.MjData(model)
    ctrl = PMPC(model, data, device=device, **params)
    ctrl.target_body = target_body
    while True:
        item = state_queue.get()
        if isinstance(item, str) and item == "STOP":
            break
        state, target = item
        data.body(ctrl.target_body).xpos[:] = [state[0], state[2], state[4]]
        data.body(ctrl.target_body).cvel[3:6] = [state[1], state[3], state[5]]
        t0 = time.time()
        u_cmd, loss = ctrl.solve(target)
        solve_time = time.time() - t0
        control_queue.put((u_cmd, loss, solve_time))
