"""PPO training step of the LMPC policy: CPU restatement with torch autograd (the reference's own library).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): only tests/, smoke() and bench.py's CPU baseline may import it.
Follows LMPC/src/controller/rlmpc2.py: ``Policy`` :33-80 (actor + critic Linear-tanh stacks, orthogonal init gain
sqrt(2), ``log_std`` clamped to [log 1e-2, log 2]), rollout-time sampling :670-699, reward :598-601 + :701-735,
``compute_gae`` :589-596, returns / advantage normalisation :783-792, the minibatch update :797-817 with
``optim.Adam(lr, weight_decay=1e-5)`` :561 and ``clip_grad_norm_(0.5)`` :816.  The module cannot be imported from the
reference itself (it pulls in casadi and mujoco at import time), hence this restatement; torch is the same library
the reference trains with, so autograd and Adam are the reference's arithmetic, not a re-derivation.

The reference's ``buf.add(obs, action, logp, value, reward, done)`` (:744) passes value and reward in the wrong order
for ``RolloutBuffer.add(o, a, logp, r, v, done)`` (:93); as SURVEY 8f.4 says, that bug is not restated.
"""
import math

import numpy as np
import torch
from torch import nn


class Policy(nn.Module):
    """rlmpc2.py:33-80; state_dict keys equal the reference's (mean_net.{0,2,4}, value_net.{0,2,4}, log_std)."""

    def __init__(self, obs_dim=520, act_dim=34, hidden_size=64, hidden_layers=2, policy_std_init=0.1,
                 policy_std_min=1e-2, policy_std_max=2.0, dtype=torch.float32):
        super().__init__()

        def stack(out_dim):
            layers, d = [], obs_dim
            for _ in range(hidden_layers):
                layers += [nn.Linear(d, hidden_size), nn.Tanh()]
                d = hidden_size
            layers.append(nn.Linear(d, out_dim))
            return nn.Sequential(*layers)

        self.mean_net = stack(act_dim)
        self.value_net = stack(1)
        self.log_std = nn.Parameter(torch.ones(act_dim) * math.log(policy_std_init))
        self.min_log_std, self.max_log_std = math.log(policy_std_min), math.log(policy_std_max)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.orthogonal_(m.weight, gain=math.sqrt(2))
                nn.init.constant_(m.bias, 0.0)
        self.to(dtype)

    def forward(self, obs):
        mean = self.mean_net(obs)
        std = torch.exp(torch.clamp(self.log_std, self.min_log_std, self.max_log_std))
        return mean, std, self.value_net(obs).squeeze(-1)


def make_policy(seed=3, dtype=torch.float32, **kw):
    torch.manual_seed(seed)
    return Policy(dtype=dtype, **kw)


def act(policy, obs, eps):
    """rlmpc2.py:670-699 with the draw made explicit: raw_action = mean + std * eps."""
    with torch.no_grad():
        mean, std, value = policy(obs)
        std = torch.clamp(std, min=1e-6)
        dist = torch.distributions.Normal(mean, std)
        action = mean + std * eps
        return action, dist.log_prob(action).sum(dim=-1), value, mean


def reward(state, target, control, prev_cmd, action, in_contact, episode_step, time_penalty, max_delta=0.1,
           action_scale=1.0, max_per_dim_rms=0.5, sigma_pos=0.02, sigma_vel=0.02, w_pos=60.0, w_vel=30.0, w_change=1e-3,
           w_d_ctrl=5.0, tray_limit=(0.2, 0.15), max_episode_steps=1000):
    """One instance, one step (rlmpc2.py:701-735).  Returns (reward, done, episode_step', time_penalty')."""
    delta_z = np.asarray(action, np.float32) * np.float32(max_delta * action_scale)
    norm = np.linalg.norm(delta_z)
    rms = norm / np.sqrt(len(delta_z))
    if rms > max_per_dim_rms:
        delta_z = delta_z * np.float32(max_per_dim_rms / (rms + 1e-12))
    pos_err = np.linalg.norm(np.abs(np.array([target[0], target[2]]) - np.array([state[0], state[2]])))
    vel_err = np.linalg.norm(np.array([state[1], state[3]]))
    pos_term = np.exp(-(pos_err ** 2) / (2 * sigma_pos ** 2))
    vel_term = np.exp(-(vel_err ** 2) / (2 * sigma_vel ** 2))
    r = w_pos * pos_term + w_vel * pos_term * vel_term
    r = r - w_change * np.linalg.norm(delta_z) - w_d_ctrl * np.sum(np.abs(np.asarray(control) - np.asarray(prev_cmd))) - time_penalty
    if pos_err < 0.01 and vel_err < 0.01:
        r += 20.0
    done = False
    episode_step += 1
    if abs(state[0]) > tray_limit[0] or abs(state[2]) > tray_limit[1]:
        r -= 20.0
        done = True
    if in_contact == 0.0:
        r -= 10.0
    if episode_step >= max_episode_steps:
        done = True
    if done:
        return float(r), True, 0, 0.0
    return float(r), False, episode_step, time_penalty + 1e-4


def compute_gae(rewards, values, dones, last_value, gamma, lam):
    """rlmpc2.py:589-596, python floats."""
    adv, gae = [], 0.0
    values = list(values) + [last_value]
    for step in reversed(range(len(rewards))):
        delta = rewards[step] + gamma * values[step + 1] * (1.0 - dones[step]) - values[step]
        gae = delta + gamma * lam * (1.0 - dones[step]) * gae
        adv.insert(0, gae)
    return adv


def normalise_returns(returns):
    """rlmpc2.py:785: numpy float64, population std."""
    returns = np.asarray(returns, dtype=np.float64)
    return (returns - returns.mean()) / (returns.std() + 1e-8)


def normalise_advantages(adv):
    """rlmpc2.py:792: torch float32, sample std."""
    a = torch.as_tensor(np.asarray(adv), dtype=torch.float32)
    return (a - a.mean()) / (a.std() + 1e-8)


def make_optimizer(policy, lr=3e-4, weight_decay=1e-5):
    return torch.optim.Adam(policy.parameters(), lr=lr, weight_decay=weight_decay)


def loss_terms(policy, obs, act_mb, old_logp, adv, ret, clip_eps=0.2, vf_coef=0.25, ent_coef=0.01):
    mean, std, val = policy(obs)
    std = torch.clamp(std, min=1e-6)
    dist = torch.distributions.Normal(mean, std)
    logp = dist.log_prob(act_mb).sum(dim=-1)
    ratio = torch.exp(logp - old_logp)
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1.0 - clip_eps, 1.0 + clip_eps) * adv
    policy_loss = -torch.min(surr1, surr2).mean()
    value_loss = nn.functional.mse_loss(val, ret)
    entropy = dist.entropy().sum(dim=-1).mean()
    return policy_loss + vf_coef * value_loss - ent_coef * entropy, policy_loss, value_loss, entropy


def minibatch_step(policy, optimizer, obs, act_mb, old_logp, adv, ret, clip_eps=0.2, vf_coef=0.25, ent_coef=0.01,
                   max_grad_norm=0.5, apply=True):
    """rlmpc2.py:801-817.  Returns (policy_loss, value_loss, entropy, grad_norm_before_clipping)."""
    loss, pl, vl, ent = loss_terms(policy, obs, act_mb, old_logp, adv, ret, clip_eps, vf_coef, ent_coef)
    optimizer.zero_grad()
    loss.backward()
    if not apply:
        gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in policy.parameters()))
        return float(pl.detach()), float(vl.detach()), float(ent.detach()), float(gn)
    gn = nn.utils.clip_grad_norm_(policy.parameters(), max_norm=max_grad_norm)
    optimizer.step()
    return float(pl.detach()), float(vl.detach()), float(ent.detach()), float(gn)
