"""Device-resident PMPC closed-loop episodes on the surrogate plant, with the reference logger's metrics.

Replaces, for B instances at once, the while-loop of PMPC/main.py:90-125 (solve every simulated step, apply the first
tilt command) with MuJoCo swapped for the surrogate plant of SURVEY 8(d)/8(f).1.  Two launches per simulated step,
no host synchronisation inside the episode.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .config import pmpc_cfg
from .engine import NMPCEngine


class PMPCEpisodes:
    """``warm_start=False`` is the reference: every PMPC solve starts cold from ``tile(state)`` / zeros (mpc_3d.py:123).
    ``warm_start=True`` reuses the previous plan as the initial point and starts the barrier of every solve after the
    first at ``warm_mu`` -- the same optimum to the solver tolerance in roughly half the iterations."""

    def __init__(self, state, target, params, mu_plant=None, coulomb=None, device=0, tol=0.01, warm_start=False,
                 warm_mu=1e-4, **cfg_kw):
        import torch
        if not torch.cuda.is_available():
            raise _lib.DartError("dart_b200 episodes need a CUDA device (no CPU fallback)")
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.cfg = pmpc_cfg(**cfg_kw)
        self.engine = NMPCEngine(self.cfg, device=device)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)
        self.B = state.shape[0]
        self.state, self.target, self.params = t(state), t(target), t(params)
        self.mu_plant = t(params[:, 3] if mu_plant is None else mu_plant)
        self.coulomb = None if coulomb is None else t(coulomb)
        f64 = torch.float64
        self.u0 = torch.zeros((self.B, 2), dtype=f64, device=self.dev)
        self.J = torch.empty((self.B,), dtype=f64, device=self.dev)
        self.status = torch.empty((self.B,), dtype=torch.int32, device=self.dev)
        self.iters = torch.empty((self.B,), dtype=torch.int32, device=self.dev)
        self.conv_time = torch.full((self.B,), -1.0, dtype=f64, device=self.dev)
        self.effort = torch.zeros((self.B,), dtype=f64, device=self.dev)
        self.err = torch.zeros((self.B,), dtype=f64, device=self.dev)
        self.nsteps = torch.zeros((self.B,), dtype=torch.int32, device=self.dev)
        self.tol = float(tol)
        self.step_index = 0
        self.warm_start, self.warm_mu = bool(warm_start), warm_mu
        if self.warm_start:
            self.w = torch.zeros((self.B, self.engine.nw), dtype=f64, device=self.dev)
            self.w_next = torch.empty_like(self.w)
        self._graph = None
        self._counters = torch.zeros(2, dtype=torch.int64, device=self.dev)      # [iterations, solves not converged]
        self.iter_sum, self.not_converged_solves = self._counters[0], self._counters[1]

    def step(self):
        torch = self.torch
        if self.warm_start:
            if self.step_index == 1 and self.warm_mu:
                self.engine.set_mu_init(self.warm_mu)
            self.engine.solve_device(self.state, self.target, aux=self.params, warm_w=self.w if self.step_index > 0 else None,
                                     w_out=self.w_next, u0_out=self.u0, J_out=self.J, status=self.status, iters=self.iters)
            self.w.copy_(self.w_next)                 # a copy, not a pointer swap: the step may be replayed from a CUDA graph
        else:
            self.engine.solve_device(self.state, self.target, aux=self.params, u0_out=self.u0, J_out=self.J,
                                     status=self.status, iters=self.iters)
        # the plant kernel also accumulates the solve statistics (no reduction launches of the loop's own)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        check(_lib.lib().dart_pmpc_plant_step(self.B, self.cfg.Ts, self.cfg.g, p(self.mu_plant), p(self.coulomb), p(self.u0),
                                              p(self.target), p(self.state), p(self.nsteps), self.tol, p(self.conv_time),
                                              p(self.effort), p(self.err), p(self.status), p(self.iters), p(self._counters),
                                              stream), "dart_pmpc_plant_step")
        self.step_index += 1

    def run(self, steps, trace_every=0, graph=False, persistent=False):
        """Run ``steps`` simulated steps; returns the metrics dict (numpy).  ``trace_every`` > 0 records
        (state, u0, J, iters) of instance 0 every that many steps (config 1's per-step solve trace).
        ``graph=True`` captures one step (solve + plant + counters) in a CUDA graph after a warm-up step and replays
        it: the closed loop is launch-bound at small batch (six launches per simulated step).
        ``persistent=True`` runs all ``steps`` in ONE launch (dart_pmpc_episode; cold start only): the same arithmetic,
        bit-identical states and metrics, no launch gaps."""
        torch = self.torch
        trace = []
        k0 = 0
        if persistent and not trace_every and not self.warm_start and steps > 0:
            p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
            counters = torch.zeros(2, dtype=torch.int64, device=self.dev)
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            with torch.cuda.device(self.dev):
                check(_lib.lib().dart_pmpc_episode(self.engine._h, self.B, int(steps), p(self.state), p(self.target), p(self.params),
                                                   p(self.mu_plant), p(self.coulomb), self.tol, p(self.nsteps), p(self.conv_time),
                                                   p(self.effort), p(self.err), p(self.u0), p(self.J), p(self.status), p(self.iters),
                                                   p(counters), stream), "dart_pmpc_episode")
            self._counters += counters
            self.step_index += int(steps)
        elif graph and not trace_every and steps > 2:
            if self._graph is None:
                self.step()                               # warm-up: one real step outside the capture
                k0 = 1
                side = torch.cuda.Stream(device=self.dev)
                side.wait_stream(torch.cuda.current_stream(self.dev))
                with torch.cuda.stream(side):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        self.step()
                    self.step_index -= 1                  # the capture itself executed nothing
                torch.cuda.current_stream(self.dev).wait_stream(side)
                self._graph = g
            for k in range(k0, steps):
                self._graph.replay()
                self.step_index += 1
        else:
            for k in range(steps):
                if trace_every and k % trace_every == 0:
                    s0 = self.state[0].cpu().numpy().copy()
                self.step()
                if trace_every and k % trace_every == 0:
                    trace.append(np.concatenate([[k * self.cfg.Ts], s0, self.u0[0].cpu().numpy(), [float(self.J[0].item()), float(self.iters[0].item())]]))
        self.torch.cuda.synchronize()
        T = self.step_index * self.cfg.Ts
        ct = self.conv_time.cpu().numpy()
        final_err = np.linalg.norm((self.state[:, [0, 2]] - self.target[:, [0, 2]]).cpu().numpy(), axis=1)
        return dict(steady_state_error=final_err, convergence_time=np.where(ct < 0, T, ct), converged=ct >= 0,
                    control_effort=self.effort.cpu().numpy(), sim_time=T, solves=self.step_index * self.B,
                    not_converged_solves=int(self.not_converged_solves.item()),
                    mean_iters=float(self.iter_sum.item()) / max(1, self.step_index * self.B),
                    trace=np.array(trace) if trace else None)
