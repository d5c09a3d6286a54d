"""PPO training of the LMPC parameter-adaptation policy on the device.

Mirrors ``RLMPC._rl_worker`` in training mode (LMPC/src/controller/rlmpc2.py:536-935): ``Policy`` actor + critic
(:33-80), sampled actions (:670-699), reward (:701-735), rollout buffer, ``compute_gae`` (:589-596), normalised
returns / advantages (:783-792), ``epochs`` x minibatch updates with clipped surrogate, value MSE and entropy bonus
(:797-817), Adam with weight decay (:561), checkpoints ``{"model", "optimizer", ...}`` (:917-922).  The reference
trains ONE instance's policy from its own rollout; here B instances share one policy and the pooled rollout
[T, B] is the buffer.  All arithmetic runs in ``csrc/ppo.cu`` through the C ABI (``dart_ppo_*``); torch only
holds the buffers and draws the random numbers (standard normal draws, minibatch permutations).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

OBS_DIM, HIDDEN, ACT_DIM = 520, 64, 34
# flat parameter layout of include/dart_b200.h (dart_ppo_nparams)
_SEGMENTS = [("W1", (128, 520)), ("b1", (128,)), ("mean_net.2.weight", (64, 64)), ("mean_net.2.bias", (64,)),
             ("value_net.2.weight", (64, 64)), ("value_net.2.bias", (64,)), ("mean_net.4.weight", (34, 64)),
             ("mean_net.4.bias", (34,)), ("value_net.4.weight", (1, 64)), ("value_net.4.bias", (1,)), ("log_std", (34,))]
NPARAMS = sum(int(np.prod(s)) for _, s in _SEGMENTS)
STATE_KEYS = ["log_std"] + [f"{net}.{i}.{wb}" for net in ("mean_net", "value_net") for i in (0, 2, 4) for wb in ("weight", "bias")]


class PPOCfg(C.Structure):
    """Mirror of ``dart_ppo_cfg``."""
    _fields_ = [(k, C.c_double) for k in ("lr", "weight_decay", "beta1", "beta2", "adam_eps", "clip_eps", "vf_coef",
                                          "ent_coef", "max_grad_norm", "log_std_min", "log_std_max")]


class PPORewardCfg(C.Structure):
    """Mirror of ``dart_ppo_reward_cfg``."""
    _fields_ = [(k, C.c_double) for k in ("max_delta", "action_scale", "max_per_dim_rms", "sigma_pos", "sigma_vel", "w_pos",
                                          "w_vel", "w_change", "w_d_ctrl", "success_tol", "success_bonus", "oob_penalty",
                                          "no_contact_penalty")] + \
               [("tray_limit", C.c_double * 2), ("max_episode_steps", C.c_int32), ("time_penalty_inc", C.c_double)]


def pack_params(state_dict):
    """Reference ``Policy.state_dict()`` (torch tensors or arrays) -> flat float32 vector."""
    g = lambda k: np.asarray(state_dict[k].detach().cpu().numpy() if hasattr(state_dict[k], "detach") else state_dict[k],
                             dtype=np.float32)
    parts = {"W1": np.concatenate([g("mean_net.0.weight"), g("value_net.0.weight")], axis=0),
             "b1": np.concatenate([g("mean_net.0.bias"), g("value_net.0.bias")])}
    out = []
    for name, shape in _SEGMENTS:
        a = parts[name] if name in parts else g(name)
        if tuple(a.shape) != shape:
            raise ValueError(f"{name}: expected {shape}, got {a.shape} (only the reference architecture 520-64-64-34/1 is supported)")
        out.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(out), dtype=np.float32)


def unpack_params(flat):
    """Flat vector -> dict of numpy arrays keyed like the reference's ``Policy.state_dict()``."""
    flat = np.asarray(flat, dtype=np.float32)
    if flat.shape != (NPARAMS,):
        raise ValueError(f"expected {NPARAMS} parameters")
    seg, o = {}, 0
    for name, shape in _SEGMENTS:
        n = int(np.prod(shape))
        seg[name] = flat[o:o + n].reshape(shape).copy()
        o += n
    sd = {k: v for k, v in seg.items() if k not in ("W1", "b1")}
    sd["mean_net.0.weight"], sd["value_net.0.weight"] = seg["W1"][:64].copy(), seg["W1"][64:].copy()
    sd["mean_net.0.bias"], sd["value_net.0.bias"] = seg["b1"][:64].copy(), seg["b1"][64:].copy()
    return {k: sd[k] for k in STATE_KEYS}


def init_policy_state(seed=3, policy_std_init=0.1):
    """Fresh ``Policy(520, 34, packet)`` parameters as the reference initialises them (orthogonal, gain sqrt(2), zero bias;
    rlmpc2.py:57-69), drawn with torch on the CPU in the module's construction order."""
    import torch
    torch.manual_seed(seed)
    sd = {}
    for net, out_dim in (("mean_net", ACT_DIM), ("value_net", 1)):
        dims = [OBS_DIM, HIDDEN, HIDDEN, out_dim]
        for li, i in enumerate((0, 2, 4)):
            lin = torch.nn.Linear(dims[li], dims[li + 1])
            sd[f"{net}.{i}.weight"], sd[f"{net}.{i}.bias"] = lin.weight, lin.bias
    for k in [k for k in sd if k.endswith("weight")]:      # _init_weights walks the modules after construction
        torch.nn.init.orthogonal_(sd[k], gain=float(np.sqrt(2)))
        torch.nn.init.constant_(sd[k.replace("weight", "bias")], 0.0)
    sd["log_std"] = torch.ones(ACT_DIM) * float(np.log(policy_std_init))
    return {k: v.detach().numpy().copy() for k, v in sd.items()}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.DartError("dart_b200 PPO needs a CUDA device (no CPU fallback)")
    return torch


class PPOTrainer:
    """Device-resident PPO learner for the LMPC policy.  ``capacity`` bounds both the number of instances per ``act``
    call and the minibatch size.  Keyword hyper-parameters carry the reference packet's names and defaults
    (rlmpc2.py:202-226): lr, clip_eps, epochs, mini_batch_size, gamma, gae_lambda, vf_coef, ent_coef, weight_decay."""

    def __init__(self, capacity, state_dict=None, seed=3, device=0, lr=3e-4, weight_decay=1e-5, clip_eps=0.2, vf_coef=0.25,
                 ent_coef=0.01, max_grad_norm=0.5, epochs=8, mini_batch_size=64, gamma=0.99, gae_lambda=0.95,
                 policy_std_init=0.1, policy_std_min=1e-2, policy_std_max=2.0, reward_cfg=None):
        self._lib = _lib.lib()
        self.torch = _torch()
        self.device = int(device)
        self.dev = self.torch.device("cuda", self.device)
        self.capacity = int(capacity)
        self.epochs, self.mini_batch_size = int(epochs), int(mini_batch_size)
        self.gamma, self.gae_lambda = float(gamma), float(gae_lambda)
        cfg = PPOCfg()
        check(self._lib.dart_ppo_default_cfg(C.byref(cfg)), "dart_ppo_default_cfg")
        cfg.lr, cfg.weight_decay, cfg.clip_eps, cfg.vf_coef = lr, weight_decay, clip_eps, vf_coef
        cfg.ent_coef, cfg.max_grad_norm = ent_coef, max_grad_norm
        cfg.log_std_min, cfg.log_std_max = float(np.log(policy_std_min)), float(np.log(policy_std_max))
        self.cfg = cfg
        self.reward_cfg = PPORewardCfg()
        check(self._lib.dart_ppo_default_reward_cfg(C.byref(self.reward_cfg)), "dart_ppo_default_reward_cfg")
        for k, v in (reward_cfg or {}).items():
            if k == "tray_limit":
                self.reward_cfg.tray_limit[0], self.reward_cfg.tray_limit[1] = float(v[0]), float(v[1])
            else:
                setattr(self.reward_cfg, k, v)
        if state_dict is None:
            state_dict = init_policy_state(seed, policy_std_init)
        flat = pack_params(state_dict)
        self._h = C.c_void_p()
        check(self._lib.dart_ppo_create(C.byref(self._h), self.device, OBS_DIM, HIDDEN, ACT_DIM, self.capacity,
                                        C.c_void_p(flat.ctypes.data), C.byref(cfg)), "dart_ppo_create")
        self.stats = self.torch.zeros(4, dtype=self.torch.float32, device=self.dev)
        self._graphs, self._gae_buf, self.graph_replays, self._gbuf = {}, None, 0, None

    # ---- plumbing ----
    def _stream(self):
        # dart_ppo_* launch on the CURRENT device and reject a mismatch with the handle's: make the handle's device current
        if self.torch.cuda.current_device() != self.device:
            self.torch.cuda.set_device(self.device)
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def _chk(self, t, dtype, shape, name):
        if t.dtype != dtype or not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: need a contiguous {dtype} CUDA tensor of shape {tuple(shape)}, got {t.dtype} {tuple(t.shape)}")
        return C.c_void_p(t.data_ptr())

    @property
    def launch_count(self):
        return int(self._lib.dart_ppo_launch_count(self._h))

    # ---- rollout ----
    def act(self, obs, eps=None):
        """obs [B,520] f32 -> (action [B,34], logp [B], value [B], mean [B,34]); ``eps`` [B,34] standard normal draws
        (None = the mean action, the reference's evaluation choice at rlmpc2.py:877-896)."""
        torch, f32 = self.torch, self.torch.float32
        B = obs.shape[0]
        po = self._chk(obs, f32, (B, OBS_DIM), "obs")
        pe = self._chk(eps, f32, (B, ACT_DIM), "eps") if eps is not None else None
        action = torch.empty((B, ACT_DIM), dtype=f32, device=self.dev)
        mean = torch.empty((B, ACT_DIM), dtype=f32, device=self.dev)
        logp = torch.empty((B,), dtype=f32, device=self.dev)
        value = torch.empty((B,), dtype=f32, device=self.dev)
        p = lambda t: C.c_void_p(t.data_ptr())
        check(self._lib.dart_ppo_act(self._h, B, po, pe, p(action), p(logp), p(value), p(mean), self._stream()), "dart_ppo_act")
        return action, logp, value, mean

    def reward(self, state, target, control, prev_cmd, action, episode_step, time_penalty, in_contact=None):
        """One control step's reward/done for B instances; ``prev_cmd``, ``episode_step``, ``time_penalty`` advance in place."""
        torch = self.torch
        B = state.shape[0]
        f64, f32 = torch.float64, torch.float32
        args = [self._chk(state, f64, (B, 8), "state"), self._chk(target, f64, (B, 8), "target"),
                self._chk(control, f64, (B, 2), "control"), self._chk(prev_cmd, f64, (B, 2), "prev_cmd"),
                self._chk(action, f32, (B, ACT_DIM), "action"),
                self._chk(in_contact, f64, (B,), "in_contact") if in_contact is not None else None,
                self._chk(episode_step, torch.int32, (B,), "episode_step"), self._chk(time_penalty, f64, (B,), "time_penalty")]
        rew = torch.empty((B,), dtype=f32, device=self.dev)
        done = torch.empty((B,), dtype=f32, device=self.dev)
        check(self._lib.dart_ppo_reward(B, C.byref(self.reward_cfg), *args, C.c_void_p(rew.data_ptr()),
                                        C.c_void_p(done.data_ptr()), self._stream()), "dart_ppo_reward")
        return rew, done

    def normalize_(self, x, ddof):
        if x.dtype != self.torch.float32 or not x.is_cuda or not x.is_contiguous():
            raise ValueError("normalize_: need a contiguous float32 CUDA tensor")
        check(self._lib.dart_ppo_normalize(x.numel(), C.c_void_p(x.data_ptr()), int(ddof), self._stream()), "dart_ppo_normalize")
        return x

    # ---- learning ----
    def update_minibatch(self, obs, act, old_logp, adv, ret, idx=None, apply=True):
        """One optimiser step on rows ``idx`` (int64 CUDA tensor; None = all rows) of the pooled rollout.  Returns the device
        tensor [policy loss, value loss, entropy, gradient norm before clipping]."""
        torch, f32 = self.torch, self.torch.float32
        S = obs.shape[0]
        M = S if idx is None else idx.shape[0]
        args = [self._chk(idx, torch.int64, (M,), "idx") if idx is not None else None,
                self._chk(obs, f32, (S, OBS_DIM), "obs"), self._chk(act, f32, (S, ACT_DIM), "act"),
                self._chk(old_logp, f32, (S,), "old_logp"), self._chk(adv, f32, (S,), "adv"), self._chk(ret, f32, (S,), "ret")]
        check(self._lib.dart_ppo_update(self._h, M, *args, 1 if apply else 0, C.c_void_p(self.stats.data_ptr()), self._stream()),
              "dart_ppo_update")
        return self.stats

    def gae(self, rewards, values, dones, last_value, out=None):
        """[T,B] f32 rollouts -> (advantages, returns), both [T,B], un-normalised (``out`` = (adv, ret) buffers to reuse)."""
        torch, f32 = self.torch, self.torch.float32
        T, B = rewards.shape
        adv, ret = out if out is not None else (torch.empty_like(rewards), torch.empty_like(rewards))
        check(self._lib.dart_ppo_gae(B, T, self._chk(rewards, f32, (T, B), "rewards"), self._chk(values, f32, (T, B), "values"),
                                     self._chk(dones, f32, (T, B), "dones"), self._chk(last_value, f32, (B,), "last_value"),
                                     self.gamma, self.gae_lambda, self._chk(adv, f32, (T, B), "adv"), self._chk(ret, f32, (T, B), "ret"),
                                     self._stream()), "dart_ppo_gae")
        return adv, ret

    def update_minibatch_distributed(self, obs, act, old_logp, adv, ret, idx=None, group=None):
        """Data-parallel optimiser step: every rank computes the gradient of its own minibatch, ONE all-reduce (NCCL) sums the
        77 317-entry gradients, every rank applies the mean with the usual clip + Adam -- the replicas stay bitwise identical.
        With equal local minibatch sizes this is the step on the concatenated minibatch."""
        import torch.distributed as dist
        torch = self.torch
        self.update_minibatch(obs, act, old_logp, adv, ret, idx=idx, apply=False)
        if self._gbuf is None:
            self._gbuf = torch.empty((NPARAMS,), dtype=torch.float32, device=self.dev)
        check(self._lib.dart_ppo_export_grad(self._h, C.c_void_p(self._gbuf.data_ptr()), self._stream()), "dart_ppo_export_grad")
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self._gbuf, group=group)
        check(self._lib.dart_ppo_apply_grad(self._h, C.c_void_p(self._gbuf.data_ptr()), 1.0 / world,
                                            C.c_void_p(self.stats.data_ptr()), self._stream()), "dart_ppo_apply_grad")
        return self.stats

    def _graph_for(self, flat, mb):
        """CUDA graph of one minibatch step over the (pointer-stable) pooled rollout ``flat`` with the row indices read from a
        static buffer.  The optimiser step count lives on the device, so the captured launches are replayable."""
        torch = self.torch
        key = (mb,) + tuple(t.data_ptr() for t in flat)
        hit = self._graphs.get(key)
        if hit is not None:
            return hit
        idx_buf = torch.zeros((mb,), dtype=torch.int64, device=self.dev)
        p, m, v, step = self._get()                         # capture must not change the learner: save, warm up + capture, restore
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.update_minibatch(*flat, idx=idx_buf)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.update_minibatch(*flat, idx=idx_buf)
        torch.cuda.synchronize(self.dev)
        cp = lambda a: C.c_void_p(a.ctypes.data)
        check(self._lib.dart_ppo_set_state(self._h, cp(p), cp(m), cp(v), step), "dart_ppo_set_state")
        if len(self._graphs) >= 4:
            self._graphs.clear()
        self._graphs[key] = (g, idx_buf)
        return g, idx_buf

    def train_rollout(self, obs, act, logp, rewards, values, dones, last_value, generator=None, graph=False):
        """The reference's update block (rlmpc2.py:777-818) on a pooled rollout: obs [T,B,520], act [T,B,34], the rest [T,B],
        last_value [B].  GAE per instance, returns normalised with the population std, advantages with the sample std,
        then ``epochs`` passes over random minibatches.  ``graph=True`` replays each full minibatch step from a CUDA graph
        (one index copy + one graph launch instead of six kernel launches; same arithmetic, bitwise) -- what pays at the
        reference's minibatch of 64, where the step is launch-bound.  Returns the number of optimiser steps taken."""
        torch = self.torch
        T, B = rewards.shape
        if graph:                                           # pointer-stable advantage / return buffers
            if self._gae_buf is None or tuple(self._gae_buf[0].shape) != (T, B):
                self._gae_buf = (torch.empty_like(rewards), torch.empty_like(rewards))
            adv, ret = self.gae(rewards, values, dones, last_value, out=self._gae_buf)
        else:
            adv, ret = self.gae(rewards, values, dones, last_value)
        self.normalize_(ret, 0)
        self.normalize_(adv, 1)
        S = T * B
        flat = (obs.reshape(S, OBS_DIM), act.reshape(S, ACT_DIM), logp.reshape(S), adv.reshape(S), ret.reshape(S))
        mb = min(self.mini_batch_size, self.capacity)
        g = idx_buf = None
        if graph and S >= mb:
            g, idx_buf = self._graph_for(flat, mb)
        steps = 0
        for _ in range(self.epochs):
            perm = torch.randperm(S, device=self.dev, generator=generator)
            for start in range(0, S, mb):
                idx = perm[start:start + mb]
                if g is not None and idx.shape[0] == mb:
                    idx_buf.copy_(idx)
                    g.replay()
                    self.graph_replays += 1
                else:
                    self.update_minibatch(*flat, idx=idx)
                steps += 1
        return steps

    # ---- state ----
    def gradient(self):
        g = np.empty(NPARAMS, dtype=np.float32)
        check(self._lib.dart_ppo_get_grad(self._h, C.c_void_p(g.ctypes.data)), "dart_ppo_get_grad")
        return unpack_params(g)

    def _get(self):
        p, m, v = (np.empty(NPARAMS, dtype=np.float32) for _ in range(3))
        step = C.c_int64()
        check(self._lib.dart_ppo_get_state(self._h, C.c_void_p(p.ctypes.data), C.c_void_p(m.ctypes.data),
                                           C.c_void_p(v.ctypes.data), C.byref(step)), "dart_ppo_get_state")
        return p, m, v, int(step.value)

    def state_dict(self):
        """Parameters keyed like the reference's ``Policy.state_dict()`` (numpy float32)."""
        return unpack_params(self._get()[0])

    def set_state(self, state_dict, exp_avg=None, exp_avg_sq=None, step=0):
        """Overwrite parameters and (optionally) the Adam moments, all keyed like ``Policy.state_dict()``, and the step count."""
        arrs = [pack_params(d) if d is not None else None for d in (state_dict, exp_avg, exp_avg_sq)]
        check(self._lib.dart_ppo_set_state(self._h, *[C.c_void_p(a.ctypes.data) if a is not None else None for a in arrs],
                                           int(step)), "dart_ppo_set_state")

    def actor_weights(self):
        """[(W,b)] x 3 of ``mean_net`` -- what ``PolicyMLP`` / ``LMPCBatch`` take as ``weights``."""
        sd = self.state_dict()
        return [(sd[f"mean_net.{i}.weight"], sd[f"mean_net.{i}.bias"]) for i in (0, 2, 4)]

    def optimizer_state_dict(self):
        """The Adam state in the layout of ``torch.optim.Adam(policy.parameters(), ...).state_dict()`` -- what the reference
        stores under ``"optimizer"`` (rlmpc2.py:917-922) and what its ``optimizer.load_state_dict`` reads.  Parameter
        indices follow ``Policy.parameters()``: the module's own ``log_std`` first, then ``mean_net``, then ``value_net``
        (= ``STATE_KEYS``)."""
        import torch
        _, m, v, step = self._get()
        em, ev = unpack_params(m), unpack_params(v)
        state = {}
        if step > 0:
            for i, k in enumerate(STATE_KEYS):
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": torch.from_numpy(em[k].copy()),
                            "exp_avg_sq": torch.from_numpy(ev[k].copy())}
        group = {"lr": float(self.cfg.lr), "betas": (float(self.cfg.beta1), float(self.cfg.beta2)), "eps": float(self.cfg.adam_eps),
                 "weight_decay": float(self.cfg.weight_decay), "amsgrad": False, "maximize": False, "foreach": None,
                 "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(STATE_KEYS)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, opt):
        """Accepts torch's ``Adam.state_dict()`` layout (the reference's checkpoints) or this package's round-1 flat layout
        ``{exp_avg, exp_avg_sq, step}``.  Returns True when moments were restored."""
        import warnings
        if not opt:
            return False
        if "state" in opt and "param_groups" in opt:
            st = opt["state"]
            if not st:
                return False
            if sorted(st.keys()) != list(range(len(STATE_KEYS))):
                warnings.warn("PPOTrainer.load: optimizer state does not cover the 13 Policy parameters; Adam moments dropped")
                return False
            p = self._get()[0]
            shapes = unpack_params(p)
            g = lambda t: np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float32)
            em = {k: g(st[i]["exp_avg"]) for i, k in enumerate(STATE_KEYS)}
            ev = {k: g(st[i]["exp_avg_sq"]) for i, k in enumerate(STATE_KEYS)}
            for k in STATE_KEYS:
                if em[k].shape != shapes[k].shape:
                    warnings.warn(f"PPOTrainer.load: optimizer state {k} has shape {em[k].shape}, expected {shapes[k].shape}; Adam moments dropped")
                    return False
            steps = {int(float(st[i]["step"])) for i in st}
            if len(steps) != 1:
                warnings.warn("PPOTrainer.load: per-parameter Adam step counts differ; using the maximum")
            m, v = pack_params(em), pack_params(ev)
            check(self._lib.dart_ppo_set_state(self._h, None, C.c_void_p(m.ctypes.data), C.c_void_p(v.ctypes.data), max(steps)),
                  "dart_ppo_set_state")
            return True
        if "exp_avg" in opt and "exp_avg_sq" in opt:
            m = np.ascontiguousarray(np.asarray(opt["exp_avg"], dtype=np.float32))
            v = np.ascontiguousarray(np.asarray(opt["exp_avg_sq"], dtype=np.float32))
            check(self._lib.dart_ppo_set_state(self._h, None, C.c_void_p(m.ctypes.data), C.c_void_p(v.ctypes.data), int(opt.get("step", 0))),
                  "dart_ppo_set_state")
            return True
        warnings.warn("PPOTrainer.load: unrecognised optimizer entry; Adam moments dropped")
        return False

    def save(self, path, obs_norm=None, **extra):
        """Checkpoint in the reference's format (rlmpc2.py:917-922): ``{"model": Policy.state_dict(), "optimizer":
        Adam.state_dict(), **extra}`` -- both entries load straight into the reference's ``Policy`` / ``optim.Adam``.
        ``obs_norm`` = dict(mean, M2, count) additionally persists the Welford observation normaliser (rlmpc2.py:552-555),
        which the reference loses on restart (its evaluation run re-estimates it from scratch); readers that do not know the
        key ignore it."""
        import torch
        p = self._get()[0]
        ck = {"model": {k: torch.from_numpy(a) for k, a in unpack_params(p).items()}, "optimizer": self.optimizer_state_dict(), **extra}
        if obs_norm is not None:
            ck["obs_norm"] = {"mean": torch.as_tensor(np.asarray(obs_norm["mean"], dtype=np.float64)),
                              "M2": torch.as_tensor(np.asarray(obs_norm["M2"], dtype=np.float64)), "count": int(obs_norm["count"])}
        torch.save(ck, path)

    def load(self, path, trust_pickle=False):
        """Reads this package's and the reference's checkpoints.  The reference's files hold numpy scalars (``"return"``), which
        a safe ``weights_only`` load rejects; ``trust_pickle=True`` opts into the reference's own ``weights_only=False``
        load (only for files you trust).  Returns the remaining entries (episode, return, obs_norm, ...)."""
        import torch
        try:
            ck = torch.load(path, map_location="cpu", weights_only=True)
        except Exception:
            if not trust_pickle:
                raise
            ck = torch.load(path, map_location="cpu", weights_only=False)
        p = pack_params(ck["model"])
        check(self._lib.dart_ppo_set_state(self._h, C.c_void_p(p.ctypes.data), None, None, 0), "dart_ppo_set_state")
        self.load_optimizer_state_dict(ck.get("optimizer"))
        return {k: ck[k] for k in ck if k not in ("model", "optimizer")}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dart_ppo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LMPCTrainer:
    """Closed-loop PPO training of the shared policy over B learning-based controllers (``RLMPC._rl_worker`` with
    ``train=True``, rlmpc2.py:630-935, batched): every control step the learner samples the action that ``LMPCBatch`` applies
    to the model parameters, the reward of rlmpc2.py:701-735 is evaluated on the observed state, every ``record_every``-th
    step (the reference's 8) a transition enters the rollout, and after ``rollout_len`` transitions per instance the pooled
    [rollout_len, B] buffer goes through ``PPOTrainer.train_rollout``.  The caller owns the plant (MuJoCo in the reference;
    ``lmpc.lmpc_plant_step`` is the surrogate) and resets the instances ``step`` reports as done."""

    def __init__(self, batch, trainer, rollout_len=32, record_every=None, generator=None, graph=False, sync_reward_cfg=True):
        torch = trainer.torch
        self.batch, self.ppo = batch, trainer
        self.B, self.dev = batch.B, trainer.dev
        if trainer.capacity < self.B:
            raise ValueError("PPOTrainer capacity must cover the number of instances")
        # a transition is recorded on exactly the steps whose action is applied to the model parameters (rlmpc2.py:742:
        # both happen under the same ``timestep % 8 == 0``); two independent knobs would train on actions that never acted
        if record_every is None:
            record_every = batch.update_every
        if int(record_every) != int(batch.update_every):
            raise ValueError(f"record_every ({record_every}) must equal LMPCBatch.update_every ({batch.update_every})")
        if sync_reward_cfg:                  # the change penalty is on the delta actually applied: max_delta of the batch
            trainer.reward_cfg.max_delta = float(batch.max_delta)
        self.T, self.every, self.gen, self.graph = int(rollout_len), int(record_every), generator, bool(graph)
        f32, B, T = torch.float32, self.B, self.T
        self.buf_obs = torch.zeros((T, B, OBS_DIM), dtype=f32, device=self.dev)
        self.buf_act = torch.zeros((T, B, ACT_DIM), dtype=f32, device=self.dev)
        self.buf_logp, self.buf_rew, self.buf_val, self.buf_done = (torch.zeros((T, B), dtype=f32, device=self.dev) for _ in range(4))
        self.prev_cmd = torch.zeros((B, 2), dtype=torch.float64, device=self.dev)
        self.episode_step = torch.zeros((B,), dtype=torch.int32, device=self.dev)
        self.time_penalty = torch.zeros((B,), dtype=torch.float64, device=self.dev)
        self.k = 0                       # transitions recorded in the current rollout
        self.updates = 0                 # optimiser steps so far
        self.mean_reward = []            # per finished rollout
        self._last = None
        batch.action_source = self._act

    def _act(self, obs, action_out):
        torch = self.ppo.torch
        eps = torch.randn((self.B, ACT_DIM), dtype=torch.float32, device=self.dev, generator=self.gen)
        a, logp, val, _ = self.ppo.act(obs, eps)
        action_out.copy_(a)
        self._last = (obs, a, logp, val)

    def step(self, state, target, in_contact=None):
        """One control step: returns (u0 [B,2], reward [B], done [B]) as device tensors."""
        control = self.batch.u_prev.clone()             # views['control'] as the RL worker sees it: the last published command
        t = self.batch.timestep
        u0 = self.batch.step(state, target)
        obs, a, logp, val = self._last
        rew, done = self.ppo.reward(state, target, control, self.prev_cmd, a, self.episode_step, self.time_penalty, in_contact)
        if t % self.every == 0:
            k = self.k
            self.buf_obs[k].copy_(obs); self.buf_act[k].copy_(a); self.buf_logp[k].copy_(logp)
            self.buf_rew[k].copy_(rew); self.buf_val[k].copy_(val); self.buf_done[k].copy_(done)
            self.k += 1
            if self.k == self.T:
                self.mean_reward.append(float(self.buf_rew.mean()))
                self.updates += self.ppo.train_rollout(self.buf_obs, self.buf_act, self.buf_logp, self.buf_rew, self.buf_val,
                                                       self.buf_done, val.clone(), generator=self.gen, graph=self.graph)
                self.k = 0
                self.batch.refresh_obs_params()          # current_k = views["model_params"].copy()  (rlmpc2.py:896)
        return u0, rew, done
