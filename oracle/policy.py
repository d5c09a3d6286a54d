"""LMPC parameter-adaptation policy: actor forward pass, observation build, logit-space update.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).
Follows LMPC/src/controller/rlmpc2.py: ``Policy.mean_net`` :33-46,71-80 (Linear-tanh stack,
orthogonal init gain sqrt(2), zero bias :64-69); observation build and Welford normaliser
:641-668; parameter update :742-759; shared-memory write with smoothing and soft clip :606-616.
"""
import numpy as np


def mlp_forward(obs, weights, dtype=np.float32):
    """mean_net(obs): weights = [(W1,b1),(W2,b2),...,(Wo,bo)], W as torch stores it ([out,in]); tanh between layers."""
    h = np.asarray(obs, dtype=dtype)
    for i, (W, b) in enumerate(weights):
        h = h @ np.asarray(W, dtype=dtype).T + np.asarray(b, dtype=dtype)
        if i + 1 < len(weights):
            h = np.tanh(h)
    return h


def orthogonal_policy_weights(obs_dim=520, act_dim=34, hidden=64, layers=2, seed=3):
    """Random-init actor weights of ``Policy(obs_dim, act_dim, {})`` under ``torch.manual_seed(seed)`` (rlmpc2.py:33-69):
    actor stack, critic stack (default inits consume the generator), then orthogonal re-draws in module order.
    Returns the actor as a list of (W, b) float32; equals the reference's construction bit for bit (tests/test_ref_pin.py)."""
    import torch
    torch.manual_seed(seed)
    dims = [obs_dim] + [hidden] * layers
    actor = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(layers)] + [torch.nn.Linear(hidden, act_dim)]
    critic = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(layers)] + [torch.nn.Linear(hidden, 1)]
    for lin in actor + critic:
        torch.nn.init.orthogonal_(lin.weight, gain=float(np.sqrt(2)))
        torch.nn.init.constant_(lin.bias, 0.0)
    return [(lin.weight.detach().numpy().copy(), lin.bias.detach().numpy().copy()) for lin in actor]


class ObsNormalizer:
    """Welford running mean/variance + history, batched over instances (rlmpc2.py:552-555, 641-668)."""

    def __init__(self, B, base_dim=52, history_len=10):
        self.mean = np.zeros((B, base_dim))
        self.M2 = np.zeros((B, base_dim))
        self.count = 0
        self.hist = np.zeros((B, history_len, base_dim), dtype=np.float32)

    def push(self, state, target, control, current_k):
        base = np.concatenate([np.asarray(state, np.float32), np.asarray(target, np.float32),
                               np.asarray(control, np.float32), np.asarray(current_k, np.float32)],
                              axis=-1).astype(np.float64)
        self.count += 1
        delta = base - self.mean
        self.mean = self.mean + delta / self.count
        delta2 = base - self.mean
        self.M2 = self.M2 + delta * delta2
        var = self.M2 / (self.count - 1) if self.count > 1 else np.ones_like(self.M2) * 1e-6
        std = np.sqrt(np.maximum(var, 1e-12)).astype(np.float32)
        norm = ((base.astype(np.float32) - self.mean.astype(np.float32)) / (std + np.float32(1e-8))).astype(np.float32)
        self.hist = np.concatenate([self.hist[:, 1:], norm[:, None, :]], axis=1)
        return self.hist.reshape(self.hist.shape[0], -1).copy()


def param_update(cur_k_shm, action, k_max=2.0, max_delta=0.02, min_k=1e-2, action_scale=1.0):
    """rlmpc2.py:745-757 in float32 as the reference computes it (torch tensors of the action's dtype)."""
    k = np.asarray(cur_k_shm, dtype=np.float32)
    a = np.asarray(action, dtype=np.float32)
    frac = np.clip(k / np.float32(k_max), np.float32(min_k / k_max), np.float32(1.0 - 1e-6))
    z_prev = np.log(frac / (np.float32(1.0) - frac))
    z_new = z_prev + a * np.float32(max_delta * action_scale)
    return (np.float32(k_max) / (np.float32(1.0) + np.exp(-z_new))).astype(np.float32)


def write_params(k_new, prev, k_max=2.0, min_k=1e-2, k_ceiling_margin=None, alpha=0.5):
    """``write_params_to_shm`` (rlmpc2.py:606-616), float64."""
    if k_ceiling_margin is None:
        k_ceiling_margin = max(1e-3, 0.05 * k_max)
    k_new = np.asarray(k_new, dtype=np.float64)
    prev = np.asarray(prev, dtype=np.float64)
    smoothed = alpha * k_new + (1 - alpha) * prev
    min_v, max_v, margin = min_k, k_max - k_ceiling_margin, 1e-3
    center = (max_v + min_v) / 2
    scale = (max_v - min_v) / 2 - margin
    return center + scale * np.tanh((smoothed - center) / scale)
