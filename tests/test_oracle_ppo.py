"""CPU checks of the PPO oracle (torch restatement of rlmpc2.py's training block) and of the host-side parameter layout."""
import os

import numpy as np
import pytest

import dart_b200
from dart_b200 import ppo
from oracle import ppo as oppo

REF_CKPT = "/root/reference/LMPC/src/checkpoints/general/best_agent.pth"


def test_init_matches_reference_policy_construction():
    """init_policy_state draws what Policy.__init__ + _init_weights draw (same RNG consumption order), key for key."""
    sd = ppo.init_policy_state(seed=3)
    ref = oppo.make_policy(seed=3).state_dict()
    assert sorted(sd) == sorted(ref)
    for k in sd:
        assert np.array_equal(sd[k], ref[k].numpy()), k
    assert np.allclose(sd["log_std"], np.log(0.1))
    W = sd["mean_net.0.weight"]
    assert np.allclose(W @ W.T, 2.0 * np.eye(64), atol=1e-4)      # orthogonal rows, gain sqrt(2)


def test_pack_unpack_roundtrip_and_layout():
    sd = oppo.make_policy(seed=5).state_dict()
    flat = ppo.pack_params(sd)
    assert flat.shape == (ppo.NPARAMS,) == (77317,)
    assert np.array_equal(flat[:520], sd["mean_net.0.weight"][0].numpy())                   # actor rows first
    assert np.array_equal(flat[64 * 520:65 * 520], sd["value_net.0.weight"][0].numpy())     # then the critic's
    assert np.array_equal(flat[-34:], sd["log_std"].numpy())
    back = ppo.unpack_params(flat)
    assert list(back) == ppo.STATE_KEYS
    for k in back:
        assert np.array_equal(back[k], sd[k].numpy()), k
    with pytest.raises(ValueError):
        bad = dict(sd)
        bad["mean_net.4.weight"] = np.zeros((30, 64), np.float32)
        ppo.pack_params(bad)


def test_abi_parameter_count_and_defaults(built):
    import ctypes as C
    L = dart_b200._lib.lib()
    assert L.dart_ppo_nparams() == ppo.NPARAMS
    c = ppo.PPOCfg()
    assert L.dart_ppo_default_cfg(C.byref(c)) == 0
    assert (c.lr, c.weight_decay, c.beta1, c.beta2, c.adam_eps) == (3e-4, 1e-5, 0.9, 0.999, 1e-8)      # rlmpc2.py:207,561
    assert (c.clip_eps, c.vf_coef, c.ent_coef, c.max_grad_norm) == (0.2, 0.25, 0.01, 0.5)               # :209,218-219,816
    assert abs(c.log_std_min - np.log(1e-2)) < 1e-15 and abs(c.log_std_max - np.log(2.0)) < 1e-15
    r = ppo.PPORewardCfg()
    assert L.dart_ppo_default_reward_cfg(C.byref(r)) == 0
    assert (r.sigma_pos, r.w_pos, r.w_vel, r.w_d_ctrl, r.max_episode_steps) == (0.02, 60.0, 30.0, 5.0, 1000)
    assert list(r.tray_limit) == [0.2, 0.15] and r.time_penalty_inc == 1e-4


def test_no_device_is_an_error(built):
    import torch
    if torch.cuda.is_available():
        return
    with pytest.raises(dart_b200.DartError):
        dart_b200.PPOTrainer(capacity=8)


def test_gae_closed_form():
    rng = np.random.default_rng(0)
    T, gamma, lam = 12, 0.99, 0.95
    r, v = rng.standard_normal(T).tolist(), rng.standard_normal(T).tolist()
    last = 0.3
    adv = oppo.compute_gae(r, v, [0.0] * T, last, gamma, lam)
    vv = v + [last]
    delta = [r[t] + gamma * vv[t + 1] - vv[t] for t in range(T)]
    for t in range(T):
        ref = sum((gamma * lam) ** k * delta[t + k] for k in range(T - t))
        assert abs(adv[t] - ref) < 1e-12
    d = [0.0] * T
    d[5] = 1.0                                                     # an episode end cuts the recursion
    adv2 = oppo.compute_gae(r, v, d, last, gamma, lam)
    assert abs(adv2[5] - (r[5] - v[5])) < 1e-12
    assert adv2[6:] == adv[6:]


def test_clipped_surrogate_gradient_structure():
    """Samples whose ratio left the clip range on the side the advantage rewards contribute no policy gradient."""
    import torch
    pol = oppo.make_policy(3, dtype=torch.float64)
    g = torch.Generator().manual_seed(1)
    M = 6
    obs = torch.randn(M, 520, generator=g, dtype=torch.float64)
    a, logp, val, mean = oppo.act(pol, obs, torch.randn(M, 34, generator=g, dtype=torch.float64))
    adv = torch.tensor([1.0, 1.0, -1.0, -1.0, 1.0, -1.0], dtype=torch.float64)
    shift = torch.tensor([-0.5, 0.5, 0.5, -0.5, 0.0, 0.0], dtype=torch.float64)    # ratio = exp(-shift): 1.65, 0.61, 0.61, 1.65, 1, 1
    loss, pl, vl, ent = oppo.loss_terms(pol, obs, a, logp + shift, adv, val.detach(), vf_coef=0.0, ent_coef=0.0)
    pol.zero_grad()
    obs.requires_grad_(False)
    mean_out = pol.mean_net(obs)
    mean_out.retain_grad()
    std = torch.exp(pol.log_std.clamp(pol.min_log_std, pol.max_log_std))
    lp = torch.distributions.Normal(mean_out, std).log_prob(a).sum(-1)
    ratio = torch.exp(lp - (logp + shift))
    (-torch.min(ratio * adv, ratio.clamp(0.8, 1.2) * adv).mean()).backward()
    gm = mean_out.grad.abs().sum(dim=1)
    assert gm[0] == 0 and gm[2] == 0                               # ratio > 1.2 with A > 0, ratio < 0.8 with A < 0: clipped
    assert gm[1] > 0 and gm[3] > 0 and gm[4] > 0 and gm[5] > 0      # the pessimistic branch and the interior keep the gradient


@pytest.mark.skipif(not os.path.exists(REF_CKPT), reason="reference tree not present on this machine")
def test_reference_checkpoint_loads_into_oracle_and_packs():
    """A checkpoint the reference trained (torch.save at rlmpc2.py:917-922) has exactly the state_dict the restated Policy and
    the flat device layout expect, and the optimiser it was trained with is the one restated (Adam, weight decay 1e-5)."""
    import torch
    ck = torch.load(REF_CKPT, map_location="cpu", weights_only=False)
    pol = oppo.Policy()
    pol.load_state_dict(ck["model"])
    flat = ppo.pack_params(ck["model"])
    back = ppo.unpack_params(flat)
    for k, v in ck["model"].items():
        assert np.array_equal(back[k], v.numpy()), k
    pg = ck["optimizer"]["param_groups"][0]
    assert pg["weight_decay"] == 1e-5 and tuple(pg["betas"]) == (0.9, 0.999) and pg["eps"] == 1e-8
    assert len(ck["optimizer"]["state"]) == 13                       # every Policy parameter was being trained


def test_oracle_reproduces_golden_outputs_on_reference_trained_weights():
    """tests/golden/ppo_reference_checkpoint.npz holds the policy the REFERENCE trained (best_agent.pth) and torch outputs on seeded
    inputs; the restated Policy / loss must reproduce them on any machine (same torch arithmetic, FP32 round-off only)."""
    import importlib.util
    from tests.helpers import ROOT
    spec = importlib.util.spec_from_file_location("make_ppo_golden", os.path.join(ROOT, "tests", "golden", "make_ppo_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ppo_reference_checkpoint.npz"))
    out = mod.compute(g["params"])
    for k in ("mean", "value", "logp", "action"):
        assert np.abs(out[k] - g[k]).max() <= 2e-5 * max(1.0, np.abs(g[k]).max()), k
    assert np.abs(out["stats"] - g["stats"]).max() <= 1e-4 * max(1.0, np.abs(g["stats"]).max())
    for k in [k for k in g.files if k.startswith("gnorm/")]:
        assert abs(float(out[k]) - float(g[k])) <= 1e-4 * float(g[k]) + 1e-9, k
    if os.path.exists(REF_CKPT):                                     # the fixture's weights are the checkpoint's, bit for bit
        import torch
        ck = torch.load(REF_CKPT, map_location="cpu", weights_only=False)
        assert np.array_equal(ppo.pack_params(ck["model"]), g["params"])
