"""Drop-in for the reference's learning-based controller ``RLMPC`` and its policy, batched and device-resident.

Mirrors LMPC/src/controller/rlmpc2.py: ``Policy.mean_net`` (:33-46, :71-80), the solver worker's NLP
(:236-519), the RL worker's observation build / parameter update in evaluation mode (:641-668, :742-759,
:606-616) and the ``RLMPC`` facade (:110-226, :986-1065).  The reference runs the solver and the policy in two
unsynchronised daemon processes over shared-memory mailboxes; here one ``step`` does, in order and on the
device: observation push -> policy forward -> (every ``update_every``-th step) parameter update -> NLP solve
warm-started from the previous solution.  In evaluation the action is the policy mean; PPO training (``ppo.py``)
plugs in through ``LMPCBatch.action_source``.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check
from .config import lmpc_cfg
from .engine import NMPCEngine

OBS_DIM, HIDDEN, ACT_DIM, BASE_DIM, HISTORY = 520, 64, 34, 52, 10


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.DartError("dart_b200 LMPC needs a CUDA device (no CPU fallback)")
    return torch


def init_policy_weights(seed=3, obs_dim=OBS_DIM, act_dim=ACT_DIM, hidden=HIDDEN, layers=2):
    """Random actor weights exactly as the reference draws them: ``Policy(obs_dim, act_dim, {})`` under
    ``torch.manual_seed(seed)`` (rlmpc2.py:33-69) builds the actor stack, then the critic stack (each ``nn.Linear`` consumes the
    generator for its default init), then ``_init_weights`` re-draws every Linear orthogonally (gain sqrt(2), zero bias)
    in module order.  The same sequence of draws is made here, so the actor equals the reference's bit for bit
    (``tests/test_ref_pin.py``, ``tests/golden/ref_policy.npz``); only the actor is returned."""
    import torch
    torch.manual_seed(seed)
    dims = [obs_dim] + [hidden] * layers
    actor = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(layers)] + [torch.nn.Linear(hidden, act_dim)]
    critic = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(layers)] + [torch.nn.Linear(hidden, 1)]
    for lin in actor + critic:
        torch.nn.init.orthogonal_(lin.weight, gain=float(np.sqrt(2)))
        torch.nn.init.constant_(lin.bias, 0.0)
    return [(lin.weight.detach().numpy().copy(), lin.bias.detach().numpy().copy()) for lin in actor]


def load_checkpoint(path, trust_pickle=False):
    """A reference checkpoint (``torch.save({"model": state_dict, "optimizer": ..., ...})``, rlmpc2.py:917-922) as a dict.
    The files contain numpy scalars, so a safe ``weights_only`` load may fail; ``trust_pickle=True`` opts into
    the reference's own ``weights_only=False`` load (only for files you trust)."""
    import torch
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        if not trust_pickle:
            raise
        return torch.load(path, map_location="cpu", weights_only=False)


def load_checkpoint_weights(path, trust_pickle=False):
    """Actor weights [(W, b)] x 3 from a reference checkpoint."""
    sd = load_checkpoint(path, trust_pickle)["model"]
    return [(sd[f"mean_net.{i}.weight"].float().numpy().copy(), sd[f"mean_net.{i}.bias"].float().numpy().copy())
            for i in (0, 2, 4)]


class PolicyMLP:
    """``Policy.mean_net`` forward on the tcgen05 kernel: obs [B,520] f32 CUDA tensor -> mean [B,34] f32."""

    def __init__(self, weights=None, seed=3, device=0, precision="fp32"):
        """precision: "fp32" (default; within 2e-5 of the reference's FP32 torch forward: 3xTF32 tensor-core layer 1, FP32
        layers 2-3) or "tf32" (single-pass TF32 tensor cores: faster streaming, |error| <= 8e-3)."""
        self._lib = _lib.lib()
        self.device = int(device)
        self.weights = init_policy_weights(seed) if weights is None else weights
        flat = []
        for W, b in self.weights:
            flat += [np.ascontiguousarray(W, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)]
        if [a.shape for a in flat] != [(64, 520), (64,), (64, 64), (64,), (34, 64), (34,)]:
            raise ValueError("PolicyMLP supports the reference architecture 520 -> 64 -> 64 -> 34")
        self._h = C.c_void_p()
        check(self._lib.dart_policy_create(C.byref(self._h), self.device, OBS_DIM, HIDDEN, ACT_DIM,
                                           *[C.c_void_p(a.ctypes.data) for a in flat]), "dart_policy_create")
        if precision not in ("fp32", "tf32"):
            raise ValueError("precision: 'fp32' or 'tf32'")
        self.precision = precision
        check(self._lib.dart_policy_set_precision(self._h, 1 if precision == "tf32" else 0), "dart_policy_set_precision")

    def forward(self, obs, out=None):
        torch = _torch()
        B = obs.shape[0]
        if obs.dtype != torch.float32 or not obs.is_cuda or not obs.is_contiguous() or tuple(obs.shape) != (B, OBS_DIM):
            raise ValueError("obs: need contiguous float32 CUDA tensor [B,520]")
        if out is None:
            out = torch.empty((B, ACT_DIM), dtype=torch.float32, device=obs.device)
        stream = C.c_void_p(torch.cuda.current_stream(obs.device).cuda_stream)
        check(self._lib.dart_policy_forward(self._h, B, C.c_void_p(obs.data_ptr()), C.c_void_p(out.data_ptr()), stream),
              "dart_policy_forward")
        return out

    @property
    def launch_count(self):
        return int(self._lib.dart_policy_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dart_policy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LMPCBatch:
    """B learning-based controllers resident on one GPU (one ``step`` = 3-4 launches).

    ``warm_start``: True = the previous solution as it is (the reference, rlmpc2.py:492,521), ``"shift"`` = the
    previous plan advanced by one stage (last stage repeated), False = cold.  ``plan_fallback``: when a solve ends
    without a usable plan (status other than converged / acceptable) the command is the next entry of the last good
    plan and that plan stays the warm start -- the reference facade's plan shift (rlmpc2.py:1013-1018), which there
    covers "the solver process has nothing new yet".  Solver options (``tol``, ``acceptable_tol``, ``acceptable_iter``,
    ``max_iter``; the reference's are ``config.LMPC_REFERENCE_SOLVER_OPTIONS``) pass through ``cfg_kw``.  ``warm_mu``: initial
    barrier parameter of the warm-started solves (every step after the first) -- with the predictor-corrector steps LMPC runs
    by default it is the scale of their initial multipliers, z = warm_mu / slack; measured on the config-4 loop
    (tools/closed_loop_strategies.py, ms per step / iterations per solve): 1e-3 3.23 / 4.21, 1e-4 3.39 / 4.39, 1e-2 3.51 / 4.68;
    under the monotone schedule 1e-4 3.39 / 5.55, 1e-6 3.27 / 5.38; None keeps the cold-start default."""

    def __init__(self, B, pvec0, weights=None, seed=3, device=0, max_param_abs=2.0, max_delta_abs=0.02, min_k=1e-2,
                 k_ceiling_margin=None, shm_smooth_alpha=0.5, update_every=8, warm_start=True, plan_fallback=False,
                 warm_mu=1e-3, dual_warm=False, obs_k0=None, live_params_in_obs=False, eval_std=None, generator=None, **cfg_kw):
        torch = _torch()
        self.torch = torch
        self.B = int(B)
        self.dev = torch.device("cuda", device)
        self.cfg = lmpc_cfg(**cfg_kw)
        self.engine = NMPCEngine(self.cfg, device=device)
        self.policy = PolicyMLP(weights, seed, device)
        self.k_max, self.max_delta, self.min_k = float(max_param_abs), float(max_delta_abs), float(min_k)
        self.margin = max(1e-3, 0.05 * self.k_max) if k_ceiling_margin is None else float(k_ceiling_margin)
        self.alpha, self.update_every = float(shm_smooth_alpha), int(update_every)
        f64, f32 = torch.float64, torch.float32
        self.aux = torch.zeros((B, 36), dtype=f64, device=self.dev)               # [u_prev(2), pvec(34)]
        self.aux[:, 2:] = torch.from_numpy(np.ascontiguousarray(pvec0, dtype=np.float64)).to(self.dev)
        # ``current_k`` of the observation (rlmpc2.py:649) is the RL worker's LOCAL copy: initialised at :618-623 and
        # refreshed only after a PPO rollout update (:896) -- not the live shared-memory parameters, which change every
        # 8th step.  Under the Welford normaliser that slice is therefore exactly 0 between updates, and a reference-
        # trained checkpoint expects that.  ``live_params_in_obs=True`` feeds the live parameters instead.
        self.live_params_in_obs = bool(live_params_in_obs)
        k_obs = pvec0 if obs_k0 is None else obs_k0
        self.obs_k = torch.from_numpy(np.array(np.broadcast_to(k_obs, (B, ACT_DIM)), dtype=np.float64)).to(self.dev).contiguous()
        # evaluation action: the policy mean (deterministic; what the reference applies after each update, :877-896).  The
        # reference's worker applies a *sample* N(mean, std) on every step even in evaluation (:678); eval_std=0.1
        # (exp(log_std) of its checkpoints) with a torch generator reproduces that behaviour.
        self.eval_std = None if eval_std is None else float(eval_std)
        self.generator = generator
        self.u_prev = torch.zeros((B, 2), dtype=f64, device=self.dev)              # views['control'] of the reference
        self.obs = [torch.zeros((B, OBS_DIM), dtype=f32, device=self.dev) for _ in range(2)]
        self.mean = torch.zeros((B, BASE_DIM), dtype=f64, device=self.dev)
        self.M2 = torch.zeros((B, BASE_DIM), dtype=f64, device=self.dev)
        self.action = torch.empty((B, ACT_DIM), dtype=f32, device=self.dev)
        self.w = torch.zeros((B, self.engine.nw), dtype=f64, device=self.dev)     # reference: w0 = zeros (rlmpc2.py:492)
        self.w_next = torch.empty_like(self.w)
        self.u0 = torch.empty((B, 2), dtype=f64, device=self.dev)
        self.J = torch.empty((B,), dtype=f64, device=self.dev)
        self.status = torch.empty((B,), dtype=torch.int32, device=self.dev)
        self.iters = torch.empty((B,), dtype=torch.int32, device=self.dev)
        self.timestep = 0
        self.count = 0
        self.action_source = None
        self.warm_start = warm_start
        self.warm_mu = warm_mu if warm_start else None
        self.dual = None
        if dual_warm and warm_start:             # dart_set_dual_state: previous slacks and multipliers, barrier from 1e-6
            self.dual = torch.zeros((B, self.engine.ndual), dtype=f64, device=self.dev)
            self.engine.set_dual_state(self.dual)
            self.warm_mu = 1e-6
        self.plan_fallback = bool(plan_fallback)
        N = self.cfg.N
        self._nxw = 8 * (N + 1)
        if self.plan_fallback:
            self.plan_U = torch.zeros((B, N, 2), dtype=f64, device=self.dev)
            self.plan_pos = torch.zeros((B,), dtype=torch.int64, device=self.dev)
            self.n_fallback = torch.zeros((), dtype=torch.int64, device=self.dev)     # solves replaced by a plan shift
            self.have_plan = torch.zeros((B,), dtype=torch.bool, device=self.dev)

    def _shifted(self, w):
        """Previous plan advanced by one stage: x_k <- x_{k+1}, u_k <- u_{k+1}, last stage repeated."""
        torch, N = self.torch, self.cfg.N
        X = w[:, :self._nxw].view(self.B, N + 1, 8)
        U = w[:, self._nxw:].view(self.B, N, 2)
        return torch.cat([X[:, 1:].reshape(self.B, -1), X[:, -1], U[:, 1:].reshape(self.B, -1), U[:, -1]], dim=1)

    @property
    def pvec(self):
        return self.aux[:, 2:]

    def obs_norm_state(self):
        """Welford state of the observation normaliser (rlmpc2.py:552-555): dict(mean [B,52], M2 [B,52], count)."""
        return {"mean": self.mean.cpu().numpy(), "M2": self.M2.cpu().numpy(), "count": int(self.count)}

    def set_obs_norm_state(self, st):
        torch = self.torch
        mean = np.broadcast_to(np.asarray(st["mean"], dtype=np.float64), (self.B, BASE_DIM))
        M2 = np.broadcast_to(np.asarray(st["M2"], dtype=np.float64), (self.B, BASE_DIM))
        self.mean.copy_(torch.from_numpy(np.ascontiguousarray(mean)))
        self.M2.copy_(torch.from_numpy(np.ascontiguousarray(M2)))
        self.count = int(st["count"])

    def refresh_obs_params(self):
        """``current_k = views["model_params"].copy()`` after a PPO rollout update (rlmpc2.py:896)."""
        self.obs_k.copy_(self.aux[:, 2:])

    @property
    def control(self):
        return self.u_prev

    def step(self, state, target, fresh=None):
        """state, target: CUDA tensors [B,8].  Returns the device tensor of first tilt commands [B,2].

        ``fresh`` (bool CUDA tensor [B], needs ``plan_fallback``): the reference's asynchronous mailbox in deterministic
        form -- where False, this call behaves as ``RLMPC.solve`` does when the solver process has published nothing new
        (rlmpc2.py:1013-1018): the command is the next entry of the last plan (held once the plan is used up, or
        ``last_control`` when there is no plan yet) and the plan stays the warm start."""
        torch, L = self.torch, _lib.lib()
        p = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        self.count += 1
        obs_in, obs_out = self.obs
        torch.cuda.nvtx.range_push("lmpc_policy")          # obs push + policy forward + parameter update
        if self.live_params_in_obs:
            k_ptr, k_stride = C.c_void_p(self.aux.data_ptr() + 16), 36
        else:
            k_ptr, k_stride = p(self.obs_k), ACT_DIM
        check(L.dart_policy_obs_push(self.B, self.count, p(state), p(target), p(self.u_prev), k_ptr,
                                     k_stride, p(self.mean), p(self.M2), p(obs_in), p(obs_out), stream), "dart_policy_obs_push")
        self.obs = [obs_out, obs_in]
        if self.action_source is None:
            self.policy.forward(obs_out, self.action)          # evaluation: the policy mean (tcgen05 kernel)
            if self.eval_std:
                self.action.add_(torch.randn(self.action.shape, dtype=self.action.dtype, device=self.dev, generator=self.generator),
                                 alpha=self.eval_std)
        else:
            self.action_source(obs_out, self.action)           # training: sampled by the PPO learner (ppo.LMPCTrainer)
        if self.update_every > 0 and self.timestep % self.update_every == 0:
            check(L.dart_policy_param_update(self.B, p(self.action), C.c_void_p(self.aux.data_ptr() + 16), 36, self.k_max,
                                             self.max_delta, self.min_k, self.margin, self.alpha, stream),
                  "dart_policy_param_update")
        torch.cuda.nvtx.range_pop()
        warm = None
        if self.warm_start:
            warm = self._shifted(self.w) if (self.warm_start == "shift" and self.timestep > 0) else self.w
        if self.timestep == 1 and self.warm_mu:
            self.engine.set_mu_init(self.warm_mu)
        self.engine.solve_device(state, target, aux=self.aux, warm_w=warm,
                                 w_out=self.w_next, u0_out=self.u0, J_out=self.J, status=self.status, iters=self.iters)
        if fresh is not None and not self.plan_fallback:
            raise ValueError("step(fresh=...) needs LMPCBatch(plan_fallback=True)")
        # one launch: plan fallback (if enabled), u_prev <- u0, aux[:, :2] <- u0
        pf = self.plan_fallback
        check(L.dart_lmpc_post_step(self.B, self.cfg.N, p(self.status), (p(fresh) if fresh is not None else None), p(self.w),
                                    p(self.w_next), p(self.u0), p(self.u_prev), p(self.aux), (p(self.plan_U) if pf else None),
                                    (p(self.plan_pos) if pf else None), (p(self.have_plan) if pf else None),
                                    (p(self.n_fallback) if pf else None), stream), "dart_lmpc_post_step")
        self.w, self.w_next = self.w_next, self.w
        self.timestep += 1
        return self.u0


def lmpc_plant_step(state, u, true_aux, Ts=0.002, out=None):
    """One surrogate-plant step (``dart_lmpc_plant_step``): the LMPC model with per-instance true parameters ``true_aux``
    [B,36] (columns 2: are the 34 parameters).  state [B,8], u [B,2] float64 CUDA tensors -> next state [B,8]."""
    torch = _torch()
    B = state.shape[0]
    for t, shape in ((state, (B, 8)), (u, (B, 2)), (true_aux, (B, 36))):
        if t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != shape:
            raise ValueError(f"lmpc_plant_step: need contiguous float64 CUDA tensors, got {t.dtype} {tuple(t.shape)} for {shape}")
    if out is None:
        out = torch.empty_like(state)
    stream = C.c_void_p(torch.cuda.current_stream(state.device).cuda_stream)
    check(_lib.lib().dart_lmpc_plant_step(B, float(Ts), C.c_void_p(true_aux.data_ptr()), C.c_void_p(u.data_ptr()),
                                          C.c_void_p(state.data_ptr()), C.c_void_p(out.data_ptr()), stream), "dart_lmpc_plant_step")
    return out


class _Event:
    """multiprocessing.Event look-alike (the reference's callers touch ``events['reset']``)."""

    def __init__(self):
        self._f = False

    def set(self):
        self._f = True

    def clear(self):
        self._f = False

    def is_set(self):
        return self._f

    def wait(self, timeout=None):
        return self._f


class RLMPC:
    """Same constructor, attributes and methods as the reference facade (rlmpc2.py:110-226, 986-1065), synchronous:
    ``solve(target)`` returns this step's first move instead of whatever the solver process last published."""

    def __init__(self, model, data, params, device=0):
        self._device = device
        self.setup(model, data, params)
        nx, nu, N = int(params["nx"]), int(params["nu"]), int(params["N"])
        self.shapes = {"state": (nx,), "state_next": (nx,), "target": (nx,), "w_opt": (nx * (N + 1) + nu * N,), "loss": (1,),
                       "control": (nu,), "model_params": (self.params_len,), "state_deriv": (nx,), "in_contact": (1,),
                       "RLstatus": (1,)}
        self.views = {k: np.zeros(s, dtype=np.float64) for k, s in self.shapes.items()}
        rng = np.random.default_rng(params.get("seed", None))
        # rlmpc2.py:143 initial mailbox content, then the RL worker's own initialisation (:618-623)
        k_max = params.get("max_param_abs", 0.5)
        min_k = params.get("min_k", 1e-2)
        margin = params.get("k_ceiling_margin", max(1e-3, 0.05 * k_max))
        base = params.get("init_k_frac", 0.5) * k_max
        jitter = rng.uniform(-params.get("init_k_jitter", 0.05), params.get("init_k_jitter", 0.05), size=self.params_len) * k_max
        k0 = np.clip(np.full(self.params_len, base) + jitter, min_k, k_max - margin)
        self.last_control = np.array([0.0, 0.0])
        self.events = {k: _Event() for k in ("state_ready", "ctrl_ready", "data_ready", "terminate", "reset")}
        weights = obs_norm = None
        ck = os.path.join(self.checkpoint_dir, "best_agent.pth")
        if not params.get("train", True) and os.path.exists(ck):
            ckd = load_checkpoint(ck, trust_pickle=params.get("trust_checkpoint_pickle", False))
            sd = ckd["model"]
            weights = [(sd[f"mean_net.{i}.weight"].float().numpy().copy(), sd[f"mean_net.{i}.bias"].float().numpy().copy())
                       for i in (0, 2, 4)]
            obs_norm = ckd.get("obs_norm")          # only this package's checkpoints carry it (see PPOTrainer.save)
        # the worker's start-up write (:623): smoothed against the mailbox's initial content (:143, uniform(0, k_max/2)), soft-clipped
        alpha = params.get("shm_smooth_alpha", 0.5)
        prev = rng.uniform(0, k_max / 2, size=self.params_len)
        lo_k, hi_k = min_k, k_max - margin
        center, scale = (hi_k + lo_k) / 2, (hi_k - lo_k) / 2 - 1e-3
        k_shm = center + scale * np.tanh((alpha * k0 + (1 - alpha) * prev - center) / scale)
        self._batch = LMPCBatch(1, k_shm[None, :], obs_k0=k0[None, :], weights=weights, seed=params.get("policy_seed", 3), device=device,
                                max_param_abs=k_max, max_delta_abs=params.get("max_delta_abs", 0.1), min_k=min_k,
                                k_ceiling_margin=margin, shm_smooth_alpha=params.get("shm_smooth_alpha", 0.5),
                                Ts=params["Ts"], N=N, Q=params["Q"], Qt=params["Qt"], R=params["R"],
                                u_bounds=params["u_bounds"], plan_fallback=True,
                                warm_start=params.get("warm_start", True), **dict(params.get("solver_options", {})))
        self.views["model_params"][:] = k_shm
        if obs_norm is not None and params.get("restore_obs_norm", True):
            self._batch.set_obs_norm_state({"mean": np.asarray(obs_norm["mean"]), "M2": np.asarray(obs_norm["M2"]), "count": obs_norm["count"]})
        self.loss = np.zeros(1)
        # Training mode (rlmpc2.py:563-578, train=True is the packet default): the RL worker's PPO learner runs on the device and
        # the action applied to the model parameters is its sample, not the mean.  With train=False and no checkpoint the
        # reference silently switches to training; here that case evaluates the freshly initialised policy (deterministic).
        self.training = bool(params.get("train", True))
        self._trainer = self._ppo = None
        self.episode_return, self.episode_count, self.best_return = 0.0, 0, -float("inf")
        self.views["in_contact"][:] = 1.0
        if self.training:
            from .ppo import LMPCTrainer, PPOTrainer
            torch = _torch()
            pk = lambda k, d: params.get(k, d)
            tl = pk("tray_limit", [0.2, 0.15])
            self._ppo = PPOTrainer(capacity=max(1, int(pk("mini_batch_size", 64))), seed=pk("policy_seed", 3), device=device,
                                   lr=pk("lr", pk("learning_rate", 3e-4)), weight_decay=pk("weight_decay", 1e-5),
                                   clip_eps=pk("clip_eps", 0.2), vf_coef=pk("vf_coef", 0.25), ent_coef=pk("ent_coef", 0.01),
                                   epochs=pk("epochs", 8), mini_batch_size=pk("mini_batch_size", 64), gamma=pk("gamma", 0.99),
                                   gae_lambda=pk("gae_lambda", 0.95), policy_std_init=pk("policy_std_init", 0.1),
                                   reward_cfg=dict(max_delta=pk("max_delta_abs", 0.1), action_scale=pk("action_scale", 1.0),
                                                   max_per_dim_rms=pk("max_per_dim_rms", 0.5), w_pos=pk("w_pos", 1.0), w_vel=pk("w_vel", 0.1),
                                                   w_change=pk("w_change", 1e-3), w_d_ctrl=pk("w_d_ctrl", 5.0), tray_limit=tl,
                                                   max_episode_steps=int(pk("max_episode_steps", 10000))))
            gen = torch.Generator(device=self._batch.dev)
            gen.manual_seed(int(pk("seed", 0) or 0))
            self._trainer = LMPCTrainer(self._batch, self._ppo, rollout_len=int(pk("rollout_len", 2048)), generator=gen)

    def rebind(self, model, data):
        self.setup(model, data, self.params)

    def setup(self, model, data, params):
        self.model, self.data, self.params = model, data, params
        self.mu = params.get("mu", 0.2)
        self.body_name = params.get("body_name", "cube2")
        self.params_len = 34
        here = os.path.dirname(os.path.abspath(__file__))
        self.checkpoint_dir = params.get("checkpoint_dir") or os.path.join(here, "checkpoints", "unnamed")

    def get_state(self):
        """rlmpc2.py:1034-1042: [px, vx, py, vy, theta_x, omega_x, theta_y, omega_y]."""
        from scipy.spatial.transform import Rotation as Rot
        b = self.data.body(self.params.get("body_name", self.body_name))
        theta = Rot.from_matrix(np.asarray(b.xmat).reshape(3, 3)).as_euler("xyz", degrees=False)[:2]
        return np.array([b.xpos[0], b.cvel[3], b.xpos[1], b.cvel[4], theta[0], b.cvel[0], theta[1], b.cvel[1]])

    def measure(self):
        state = self.get_state()
        self.views["state"][:] = state
        return {"state": state, "state_deriv": self.views["state_deriv"].copy(), "contact": float(self.views["in_contact"][0])}

    def solve(self, target):
        torch = _torch()
        state = self.get_state()
        self.views["state"][:] = state
        self.views["target"][:] = target
        dev = self._batch.dev
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)[None, :]).to(dev)
        if self._trainer is not None:
            u0_t, rew, done = self._trainer.step(t(state), t(target), in_contact=torch.from_numpy(self.views["in_contact"].copy()).to(dev))
            u0 = u0_t.cpu().numpy()[0]
            self.episode_return += float(rew[0])
            if bool(done[0]):
                self._episode_end()
        else:
            u0 = self._batch.step(t(state), t(target)).cpu().numpy()[0]
        self.last_control = u0.astype(np.float64)
        self.loss = self._batch.J.cpu().numpy().copy()
        self.views["w_opt"][:] = self._batch.w.cpu().numpy()[0]
        self.views["loss"][:] = self.loss
        self.views["control"][:] = self.last_control
        self.views["model_params"][:] = self._batch.pvec.cpu().numpy()[0]
        return self.last_control.copy(), self.loss

    def _episode_end(self):
        """rlmpc2.py:900-926: count the episode, checkpoint (latest always, best on a new best return), ask the caller for a reset."""
        self.episode_count += 1
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        meta = {"episode": self.episode_count, "return": self.episode_return, "episode number": self.episode_count,
                "obs_norm": self._batch.obs_norm_state()}
        if self.episode_return > self.best_return:
            self.best_return = self.episode_return
            self._ppo.save(os.path.join(self.checkpoint_dir, "best_agent.pth"), **meta)
        self._ppo.save(os.path.join(self.checkpoint_dir, "latest_agent.pth"), **meta)
        self.events["reset"].set()
        self.episode_return = 0.0

    def close(self):
        self.events["terminate"].set()
        if getattr(self, "_ppo", None) is not None:
            self._ppo.close()
            self._ppo = self._trainer = None
        if getattr(self, "_batch", None) is not None:
            self._batch.engine.close()
            self._batch.policy.close()
            self._batch = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
