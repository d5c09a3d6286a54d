"""Drop-ins for the reference's adaptive controller: ``RLS`` and ``AdaptiveNPMPCSmooth``, plus the
device-resident batched closed-loop stepper that replaces one ``rob_ctrl.py`` loop body per instance.

Mirrors RMPC/dev_dual/controller/np_mpc_adaptive_with_linear_regressor.py:10-30 (``RLS``), :35-222
(``AdaptiveNPMPCSmooth``: constructor signature, ``.N .nx .v_eps .gz .w0``, ``solve(x0, u_prev, theta_hat,
Rref_flat) -> (u0, loss)``, static ``build_ref_traj``) and the caller glue of
RMPC/dev_dual/rob_ctrl.py:281-288, 335-352.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .config import rmpc_cfg
from .engine import NMPCEngine


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.DartError("dart_b200 RMPC needs a CUDA device (no CPU fallback)")
    return torch


def rls_update_device(theta, P, phi, y, lam):
    """In-place batched RLS update on CUDA tensors: theta [B,E,7], P [B,E,7,7], phi [B,7], y [B,E]."""
    torch = _torch()
    B, E = theta.shape[0], theta.shape[1]
    for t, shp in ((theta, (B, E, 7)), (P, (B, E, 7, 7)), (phi, (B, 7)), (y, (B, E))):
        if tuple(t.shape) != shp or t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"rls_update_device: need contiguous float64 CUDA tensor of shape {shp}")
    stream = C.c_void_p(torch.cuda.current_stream(theta.device).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    check(_lib.lib().dart_rls_update(B, E, p(theta), p(P), p(phi), p(y), float(lam), stream), "dart_rls_update")


def rmpc_plant_step_device(x, u, mu, c, Ts=0.002, gz=-9.81, out=None):
    """Surrogate plant of BASELINE config 3 on the device (``dart_rmpc_plant_step``; host twin: workloads.rmpc_plant_step).
    x [B,4], u [B,2], mu [B], c [B] float64 CUDA tensors -> next state [B,4]."""
    torch = _torch()
    B = x.shape[0]
    for t, shp in ((x, (B, 4)), (u, (B, 2)), (mu, (B,)), (c, (B,))):
        if tuple(t.shape) != shp or t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"rmpc_plant_step_device: need contiguous float64 CUDA tensor of shape {shp}")
    if out is None:
        out = torch.empty_like(x)
    stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    check(_lib.lib().dart_rmpc_plant_step(B, float(Ts), float(gz), p(mu), p(c), p(u), p(x), p(out), stream), "dart_rmpc_plant_step")
    return out


class RLS:
    """Scalar-output recursive least squares with forgetting; state lives on the GPU.

    Same constructor and methods as the reference class (np_mpc_adaptive...py:10-30)."""

    def __init__(self, p, theta0=None, P0=1e3, lam=0.995, device=0):
        if p != 7:
            raise ValueError("the CUDA RLS kernel is built for the reference's p = 7 regressor")
        torch = _torch()
        self.p = p
        self.lam = float(lam)
        self._dev = torch.device("cuda", device)
        th = np.zeros(p) if theta0 is None else np.asarray(theta0, dtype=float)
        self._theta = torch.from_numpy(th.reshape(1, 1, p).copy()).to(self._dev)
        self._P = (torch.eye(p, dtype=torch.float64, device=self._dev) * float(P0)).reshape(1, 1, p, p).contiguous()

    def update(self, phi, y):
        torch = _torch()
        phi = torch.from_numpy(np.asarray(phi, dtype=float).reshape(1, -1).copy()).to(self._dev)
        yv = torch.tensor([[float(np.asarray(y).reshape(()))]], dtype=torch.float64, device=self._dev)
        rls_update_device(self._theta, self._P, phi, yv, self.lam)

    def get(self):
        return self._theta.reshape(-1).cpu().numpy().copy()

    @property
    def theta(self):
        return self.get()

    @property
    def P(self):
        return self._P.reshape(self.p, self.p).cpu().numpy().copy()


class AdaptiveNPMPCSmooth:
    def __init__(self, model, data, Ts, nx=4, nu=2, N=20, Qp=100.0, Qv=1.0, Ru=0.05, Rdu=1.0, u_bounds=(-0.4, 0.4),
                 du_bounds=(-0.05, 0.05), vmax=0.25, v_eps=0.1, target_body="cube", device=0, **solver):
        self.model = model
        self.data = data
        self.Ts = float(Ts)
        self.nx, self.nu, self.N = nx, nu, N
        self.Qp, self.Qv, self.Ru, self.Rdu = Qp, Qv, Ru, Rdu
        self.u_bounds, self.du_bounds = u_bounds, du_bounds
        self.vmax, self.v_eps = vmax, v_eps
        self.target_body = target_body
        self.gz = float(model.opt.gravity[2])          # negative (np_mpc...:56)
        self.px = self.py = 7
        self.p_total = 14
        self._engine = NMPCEngine(rmpc_cfg(Ts=self.Ts, nx=nx, nu=nu, N=N, Qp=Qp, Qv=Qv, Ru=Ru, Rdu=Rdu, u_bounds=u_bounds,
                                           du_bounds=du_bounds, vmax=vmax, v_eps=v_eps, gz=self.gz, **solver), device=device)
        self.w0 = np.zeros(nx * (N + 1) + nu * N)
        self.status = None
        self.iters = None

    def get_state(self):
        pos = self.data.body(self.target_body).xpos[:2]
        vxy = self.data.body(self.target_body).cvel[3:5]
        return np.array([pos[0], vxy[0], pos[1], vxy[1]], dtype=float)

    @staticmethod
    def build_ref_traj(x_now, r_v, target, N, nx, step_fraction=0.2):
        """np_mpc_adaptive...py:201-210 (host helper, unchanged semantics)."""
        R = np.zeros(((N + 1), nx), dtype=float)
        for i in range(N + 1):
            w = 1.0 - (1.0 - step_fraction) ** (i + 1)
            r_i = r_v + w * (target - r_v)
            R[i, :] = np.array([r_i[0], 0.0, r_i[2], 0.0])
        return R.reshape(-1)

    def solve(self, x0, u_prev, theta_hat, Rref_flat):
        """Primal warm start from the previous solution, exactly as np_mpc...:212-222 passes ``x0=self.w0``."""
        aux = np.concatenate([np.asarray(u_prev, dtype=float), np.asarray(theta_hat, dtype=float)])[None, :]
        out = self._engine.solve(np.asarray(x0, dtype=float)[None, :], np.asarray(Rref_flat, dtype=float)[None, :],
                                 aux=aux, warm_w=self.w0[None, :])
        self.w0 = out["w"][0]
        self.status = int(out["status"][0])
        self.iters = int(out["iters"][0])
        return out["u0"][0].copy(), out["J"].copy()

    def solve_batch(self, x0, u_prev, theta_hat, Rref_flat, warm_w=None, want_w=True):
        aux = np.concatenate([np.atleast_2d(u_prev), np.atleast_2d(theta_hat)], axis=1)
        return self._engine.solve(x0, Rref_flat, aux=aux, warm_w=warm_w, want_w=want_w)

    @property
    def engine(self):
        return self._engine


class RMPCBatch:
    """B adaptive controllers resident on one GPU: the loop body of rob_ctrl.py:333-352 for every instance,
    two launches per step (fused prologue, solve).  State kept on the device between steps: RLS theta/P,
    virtual reference r_v, previous state and command, primal warm start."""

    def __init__(self, B, target, x_init, device=0, rls_P0=1e3, rls_lam=0.995, dr_max=0.01, alpha_rg=0.5,
                 step_fraction=0.2, warm_start=True, warm_mu=1e-4, dual_warm=False, **cfg_kw):
        torch = _torch()
        self.torch = torch
        self.B = int(B)
        self.dev = torch.device("cuda", device)
        self.cfg = rmpc_cfg(**cfg_kw)
        self.engine = NMPCEngine(self.cfg, device=device)
        self.N = self.cfg.N
        self.lam, self.dr_max, self.alpha_rg, self.step_fraction = float(rls_lam), float(dr_max), float(alpha_rg), float(step_fraction)
        f64 = torch.float64
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)
        self.target = t(target)
        self.prev_state = t(x_init)
        self.r_v = torch.zeros((B, 4), dtype=f64, device=self.dev)
        self.theta = torch.zeros((B, 2, 7), dtype=f64, device=self.dev)
        self.P = (torch.eye(7, dtype=f64, device=self.dev) * float(rls_P0)).repeat(B, 2, 1, 1).contiguous()
        self.ref = torch.empty((B, (self.N + 1) * 4), dtype=f64, device=self.dev)
        self.aux = torch.empty((B, 16), dtype=f64, device=self.dev)
        self.w = torch.zeros((B, self.engine.nw), dtype=f64, device=self.dev)      # reference: w0 = zeros at first call
        self.w_next = torch.empty_like(self.w)
        self.u0 = torch.zeros((B, 2), dtype=f64, device=self.dev)          # also u_prev of the next step (rob_ctrl.py: u_prev = 0 at start)
        self.u_prev = self.u0
        self.J = torch.empty((B,), dtype=f64, device=self.dev)
        self.status = torch.empty((B,), dtype=torch.int32, device=self.dev)
        self.iters = torch.empty((B,), dtype=torch.int32, device=self.dev)
        self.warm_start = warm_start
        # initial barrier parameter of the warm-started solves (every step after the first); None keeps 0.1
        self.warm_mu = warm_mu if warm_start else None
        self.steps = 0
        # dual warm start (dart_set_dual_state): slacks and multipliers of the previous solve, barrier from 1e-6
        self.dual = None
        if dual_warm and warm_start:
            self.dual = torch.zeros((B, self.engine.ndual), dtype=f64, device=self.dev)
            self.engine.set_dual_state(self.dual)
            self.warm_mu = 1e-6

    def set_virtual_reference(self, r_v):
        self.r_v.copy_(self.torch.from_numpy(np.ascontiguousarray(r_v, dtype=np.float64)).to(self.dev))

    def step(self, xk):
        """xk: CUDA tensor [B,4] (current states).  Returns the device tensor of first tilt commands [B,2]."""
        torch = self.torch
        p = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        c = self.cfg
        torch.cuda.nvtx.range_push("rmpc_prologue")
        check(_lib.lib().dart_rmpc_prologue(self.B, self.N, c.Ts, c.v_eps, self.lam, self.dr_max, self.alpha_rg,
                                            self.step_fraction, p(xk), p(self.prev_state), p(self.target), p(self.u_prev),
                                            p(self.r_v), p(self.theta), p(self.P), p(self.ref), p(self.aux), stream),
              "dart_rmpc_prologue")
        torch.cuda.nvtx.range_pop()
        if self.steps == 1 and self.warm_mu:
            self.engine.set_mu_init(self.warm_mu)
        self.engine.solve_device(xk, self.ref, aux=self.aux, warm_w=self.w if self.warm_start else None,
                                 w_out=self.w_next, u0_out=self.u0, J_out=self.J, status=self.status, iters=self.iters)
        self.steps += 1
        self.w, self.w_next = self.w_next, self.w
        # no eager glue: the prologue kernel has set prev_state <- xk, and the next prologue reads u_prev from u0
        return self.u0
