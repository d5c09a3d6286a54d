"""PPO training path (csrc/ppo.cu through the C ABI) against torch autograd + torch.optim.Adam on the CPU (oracle/ppo.py).

Tolerances (FP32 arithmetic on both sides, different summation orders):
  forward mean / value, log-probability      |d| <= 2e-5 * max(1, |ref|)
  gradient (before clipping)                 |d| <= 1e-6 + 2e-4 * max|g| of the tensor (FP32 accumulation over up to 16 384 samples)
                                             and ||d||_2 <= 1e-4 ||g||_2 per tensor
  one optimiser step from the same state     98 % of the elements within 1 % of lr, 99.5 % within lr / 4, all within 2 lr (an Adam step is
                                             lr * g / (|g| + 1e-8): elements whose clipped gradient is at the 1e-8 floor turn FP32
                                             summation-order noise into a fraction of one step)
  GAE / reward / normalisation               1e-5 relative
"""
import os

import numpy as np
import pytest

import dart_b200
from oracle import ppo as oppo

pytestmark = pytest.mark.gpu


def _rollout(M, seed, policy):
    """Synthetic minibatch: observations, actions sampled near the policy, old log-probabilities off by a little, advantages, returns."""
    import torch
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(M, 520, generator=g)
    eps = torch.randn(M, 34, generator=g)
    action, logp, value, mean = oppo.act(policy, obs, eps)
    old_logp = logp + 0.3 * torch.randn(M, generator=g)          # ratios on both sides of the clip range
    adv = torch.randn(M, generator=g)
    ret = value + torch.randn(M, generator=g)
    return obs, eps, action, logp, value, mean, old_logp, adv, ret


def _perturbed_policy(seed=3):
    import torch
    pol = oppo.make_policy(seed)
    g = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for p in pol.parameters():                              # non-zero biases, log_std off its initial value
            if p.ndim == 1:
                p.add_(0.05 * torch.randn(p.shape, generator=g))
    return pol


@pytest.mark.parametrize("B", [1, 63, 200, 4096])
def test_act_matches_torch(built, B):
    import torch
    pol = _perturbed_policy()
    obs, eps, action, logp, value, mean, *_ = _rollout(B, B, pol)
    tr = dart_b200.PPOTrainer(capacity=B, state_dict=pol.state_dict())
    a, lp, v, mu = tr.act(obs.cuda(), eps.cuda())
    tol = lambda ref: 2e-5 * max(1.0, float(ref.abs().max()))
    assert (mu.cpu() - mean).abs().max() <= tol(mean)
    assert (v.cpu() - value).abs().max() <= tol(value)
    assert (a.cpu() - action).abs().max() <= tol(action)
    assert (lp.cpu() - logp).abs().max() <= tol(logp)
    a0, lp0, _, _ = tr.act(obs.cuda(), None)                       # evaluation: the mean action
    assert torch.equal(a0.cpu(), mu.cpu())
    tr.close()


@pytest.mark.parametrize("M", [1, 64, 257, 1000, 16384])
def test_gradient_matches_autograd(built, M):
    import torch
    pol = _perturbed_policy()
    obs, eps, action, logp, value, mean, old_logp, adv, ret = _rollout(M, M + 7, pol)
    opt = oppo.make_optimizer(pol)
    pl, vl, ent, gn = oppo.minibatch_step(pol, opt, obs, action, old_logp, adv, ret, apply=False)
    tr = dart_b200.PPOTrainer(capacity=M, state_dict=pol.state_dict())
    stats = tr.update_minibatch(obs.cuda(), action.cuda(), old_logp.cuda(), adv.cuda(), ret.cuda(), apply=False).cpu().numpy()
    grad = tr.gradient()
    # log-probabilities are ~30 in FP32 (ulp 2e-6, 34-term sums): the ratio exp(logp - old) carries ~3e-5 relative noise on both sides
    assert abs(stats[0] - pl) <= 1e-4 * max(1.0, abs(pl)) and abs(stats[1] - vl) <= 1e-5 * max(1.0, abs(vl))
    assert abs(stats[2] - ent) <= 1e-5 * abs(ent)
    for k, p in pol.named_parameters():
        ref = p.grad.numpy()
        assert grad[k].shape == ref.shape
        err = np.abs(grad[k] - ref).max()
        assert err <= 1e-6 + 2e-4 * np.abs(ref).max(), (k, err, np.abs(ref).max())
        rel2 = np.linalg.norm((grad[k] - ref).astype(np.float64)) / max(np.linalg.norm(ref.astype(np.float64)), 1e-30)
        assert rel2 <= 1e-4, (k, rel2)                                   # no structured error hiding under the element-wise bound
    before = tr.state_dict()
    assert all(np.array_equal(before[k], v.detach().numpy()) for k, v in pol.state_dict().items())   # apply=False leaves the parameters alone
    tr.close()


@pytest.mark.parametrize("M,steps", [(64, 6), (1000, 4)])
def test_optimizer_steps_match_torch_adam(built, M, steps):
    """Step by step from the same state (parameters, Adam moments, step count).  A free-running comparison is ill-conditioned:
    std = 0.1 amplifies parameter differences 10x into the log-probabilities, and an Adam step is lr * g / (|g| + 1e-8), so the
    0.3-0.5 % of elements whose clipped gradient is below 1e-7 turn FP32 summation-order noise into O(lr) differences."""
    import torch
    lr = 3e-4
    pol = _perturbed_policy()
    opt = oppo.make_optimizer(pol, lr=lr, weight_decay=1e-5)
    tr = dart_b200.PPOTrainer(capacity=M, state_dict=pol.state_dict(), lr=lr, weight_decay=1e-5)
    names = [k for k, _ in pol.named_parameters()]
    for s in range(steps):
        obs, eps, action, logp, value, mean, old_logp, adv, ret = _rollout(M, 1000 + s, pol)
        before = {k: v.detach().clone().numpy() for k, v in pol.state_dict().items()}
        if s > 0:
            st = opt.state_dict()["state"]
            tr.set_state(before, {k: st[i]["exp_avg"] for i, k in enumerate(names)},
                         {k: st[i]["exp_avg_sq"] for i, k in enumerate(names)}, step=s)
        pl, vl, ent, gn = oppo.minibatch_step(pol, opt, obs, action, old_logp, adv, ret)
        stats = tr.update_minibatch(obs.cuda(), action.cuda(), old_logp.cuda(), adv.cuda(), ret.cuda()).cpu().numpy()
        assert abs(stats[3] - gn) <= 5e-5 * gn, (s, stats[3], gn)
        assert abs(stats[0] - pl) <= 1e-4 * max(1.0, abs(pl)), (s, stats[0], pl)
        sd = tr.state_dict()
        for k, p in pol.state_dict().items():
            d_ref, d_gpu = p.numpy() - before[k], sd[k] - before[k]
            err = np.abs(d_gpu - d_ref)
            assert err.max() <= 1e-7 + 2 * lr, (s, k, err.max())         # a sign flip of a noise-floor gradient is the worst case
            assert np.mean(err <= 1e-7 + 0.25 * lr) >= 0.995, (s, k, np.mean(err <= 1e-7 + 0.25 * lr))
            assert np.mean(err <= 1e-7 + 1e-2 * lr) >= 0.98, (s, k, np.mean(err <= 1e-7 + 1e-2 * lr))
            assert np.abs(d_gpu).max() > 0.1 * lr, (s, k)               # the parameters did move
    tr.close()


def test_minibatch_gather_and_repeatability(built):
    import torch
    pol = _perturbed_policy()
    S, M = 3000, 512
    obs, eps, action, logp, value, mean, old_logp, adv, ret = _rollout(S, 11, pol)
    idx = torch.randperm(S, generator=torch.Generator().manual_seed(5))[:M]
    d = [t.cuda() for t in (obs, action, old_logp, adv, ret)]
    outs = []
    for rep in range(3):
        tr = dart_b200.PPOTrainer(capacity=M, state_dict=pol.state_dict())
        if rep < 2:
            tr.update_minibatch(*d, idx=idx.cuda())
        else:                                                   # gathered on the host instead
            tr.update_minibatch(*[t[idx].contiguous().cuda() for t in (obs, action, old_logp, adv, ret)])
        outs.append(tr.state_dict())
        tr.close()
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k          # bitwise repeatable
        assert np.array_equal(outs[0][k], outs[2][k]), k          # device gather == host gather


@pytest.mark.parametrize("M", [777, 4096])
def test_tensor_core_layer1_matches_simt_layer1(built, M, monkeypatch):
    """The tcgen05 layer-1 GEMMs (3xTF32, csrc/ppo_tc.cu: forward and weight gradient, minibatch gathered by the loaders,
    ragged last tile / last sample chunk) against the FP32 SIMT kernels they replace (DART_PPO_SIMT=1): same gradient to
    FP32 summation-order noise, and the tensor-core path is bitwise repeatable."""
    import torch
    pol = _perturbed_policy()
    S = M + 333
    obs, eps, action, logp, value, mean, old_logp, adv, ret = _rollout(S, 23, pol)
    idx = torch.randperm(S, generator=torch.Generator().manual_seed(9))[:M].cuda()
    d = [t.cuda() for t in (obs, action, old_logp, adv, ret)]
    grads = []
    for mode in ("tc", "tc", "simt"):
        if mode == "simt":
            monkeypatch.setenv("DART_PPO_SIMT", "1")
        else:
            monkeypatch.delenv("DART_PPO_SIMT", raising=False)
        tr = dart_b200.PPOTrainer(capacity=M, state_dict=pol.state_dict())
        tr.update_minibatch(*d, idx=idx, apply=False)
        grads.append(tr.gradient())
        tr.close()
    for k in grads[0]:
        assert np.array_equal(grads[0][k], grads[1][k]), k
        ref = grads[2][k]
        # the actor's gradient carries ~2e-5 of FP32 noise in either path (std = 0.1 amplifies the forward's rounding ~100x)
        assert np.abs(grads[0][k] - ref).max() <= 1e-7 + 1e-4 * np.abs(ref).max(), (k, np.abs(grads[0][k] - ref).max(), np.abs(ref).max())


def test_gae_reward_normalise(built):
    import torch
    rng = np.random.default_rng(4)
    T, B = 37, 301
    rew = rng.standard_normal((T, B)).astype(np.float32)
    val = rng.standard_normal((T, B)).astype(np.float32)
    done = (rng.random((T, B)) < 0.05).astype(np.float32)
    last = rng.standard_normal(B).astype(np.float32)
    tr = dart_b200.PPOTrainer(capacity=64, gamma=0.99, gae_lambda=0.95)
    adv, ret = tr.gae(*[torch.from_numpy(a).cuda() for a in (rew, val, done, last)])
    for b in (0, 17, 300):
        ref = np.array(oppo.compute_gae(rew[:, b].astype(float).tolist(), val[:, b].astype(float).tolist(),
                                        done[:, b].astype(float).tolist(), float(last[b]), 0.99, 0.95))
        assert np.abs(adv[:, b].cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
        assert np.abs(ret[:, b].cpu().numpy() - (ref + val[:, b])).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    a_np, r_np = adv.cpu().numpy().reshape(-1).copy(), ret.cpu().numpy().reshape(-1).copy()
    tr.normalize_(ret, 0)
    tr.normalize_(adv, 1)
    assert np.abs(ret.cpu().numpy().reshape(-1) - oppo.normalise_returns(r_np)).max() <= 1e-5
    assert np.abs(adv.cpu().numpy().reshape(-1) - oppo.normalise_advantages(a_np).numpy()).max() <= 1e-5

    # reward: near the target, far away, out of bounds, lost contact, episode cap
    Bn = 64
    state = np.zeros((Bn, 8)); target = np.zeros((Bn, 8))
    state[:, 0] = rng.uniform(-0.25, 0.25, Bn); state[:, 2] = rng.uniform(-0.18, 0.18, Bn)
    state[:, 1] = rng.uniform(-0.1, 0.1, Bn); state[:, 3] = rng.uniform(-0.1, 0.1, Bn)
    target[:, 0] = rng.uniform(-0.1, 0.1, Bn); target[:, 2] = rng.uniform(-0.1, 0.1, Bn)
    state[:8, :4] = target[:8, :4] + 1e-3                           # success bonus
    control = rng.uniform(-0.4, 0.4, (Bn, 2)); prev = rng.uniform(-0.4, 0.4, (Bn, 2))
    action = (rng.standard_normal((Bn, 34)) * np.where(np.arange(Bn)[:, None] % 4 == 0, 40.0, 1.0)).astype(np.float32)   # some damped
    contact = (rng.random(Bn) > 0.2).astype(np.float64)
    step = rng.integers(0, 1000, Bn).astype(np.int32); step[-3:] = 999
    tpen = rng.uniform(0, 0.1, Bn)
    t = lambda a: torch.from_numpy(a.copy()).cuda()
    prev_d, step_d, tpen_d = t(prev), t(step), t(tpen)
    r, dn = tr.reward(t(state), t(target), t(control), prev_d, t(action), step_d, tpen_d, in_contact=t(contact))
    for b in range(Bn):
        rr, dd, s2, tp2 = oppo.reward(state[b], target[b], control[b], prev[b], action[b], contact[b], int(step[b]), float(tpen[b]))
        assert abs(float(r[b]) - rr) <= 1e-5 * max(1.0, abs(rr)), b
        assert bool(dn[b]) == dd and int(step_d[b]) == s2 and abs(float(tpen_d[b]) - tp2) < 1e-12, b
    assert torch.equal(prev_d.cpu(), torch.from_numpy(control))
    tr.close()


def test_train_rollout_improves_surrogate_and_checkpoint_roundtrip(built, tmp_path):
    """End to end on a synthetic pooled rollout: after the reference's epochs x minibatches the value loss has dropped, the
    checkpoint loads into the reference's Policy class (oracle restatement, same state_dict keys) and back into a trainer."""
    import torch
    T, B = 16, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    tr = dart_b200.PPOTrainer(capacity=1024, epochs=4, mini_batch_size=1024, lr=1e-3)
    obs = torch.randn(T, B, 520, device="cuda", generator=g)
    eps = torch.randn(T, B, 34, device="cuda", generator=g)
    act, logp, val, _ = tr.act(obs.reshape(-1, 520)[:1024], eps.reshape(-1, 34)[:1024])
    acts, logps, vals = [], [], []
    for s in range(0, T * B, 1024):
        a, lp, v, _ = tr.act(obs.reshape(-1, 520)[s:s + 1024], eps.reshape(-1, 34)[s:s + 1024])
        acts.append(a); logps.append(lp); vals.append(v)
    act, logp, val = torch.cat(acts).view(T, B, 34), torch.cat(logps).view(T, B), torch.cat(vals).view(T, B)
    rew = obs[:, :, 0].contiguous() + 0.1 * torch.randn(T, B, device="cuda", generator=g)      # learnable from the observation
    done = torch.zeros(T, B, device="cuda")
    last = torch.zeros(B, device="cuda")
    adv, ret = tr.gae(rew, val, done, last)
    tr.normalize_(ret, 0); tr.normalize_(adv, 1)
    flat = (obs.reshape(-1, 520), act.reshape(-1, 34), logp.reshape(-1), adv.reshape(-1), ret.reshape(-1))
    v0 = float(tr.update_minibatch(*[f[:1024].contiguous() for f in flat], apply=False)[1])
    steps = tr.train_rollout(obs, act, logp, rew, val, done, last, generator=g)
    assert steps == 4 * (T * B // 1024) and tr.launch_count >= steps * 5   # gather + 5 kernels per step
    v1 = float(tr.update_minibatch(*[f[:1024].contiguous() for f in flat], apply=False)[1])
    assert v1 < 0.8 * v0, (v0, v1)
    path = str(tmp_path / "best_agent.pth")
    tr.save(path, episode=3)
    ck = torch.load(path, map_location="cpu", weights_only=True)
    pol = oppo.Policy()
    pol.load_state_dict(ck["model"])                               # the reference's keys and shapes
    mean_ref = oppo.act(pol, obs[0, :8].cpu(), torch.zeros(8, 34))[3]
    tr2 = dart_b200.PPOTrainer(capacity=64)
    extra = tr2.load(path)
    assert extra["episode"] == 3
    mean2 = tr2.act(obs[0, :8].contiguous(), None)[3].cpu()
    assert (mean2 - mean_ref).abs().max() <= 2e-5
    sd1, sd2 = tr.state_dict(), tr2.state_dict()
    assert all(np.array_equal(sd1[k], sd2[k]) for k in sd1)
    pm = dart_b200.PolicyMLP(tr.actor_weights())                   # the trained actor feeds the tcgen05 inference kernel
    assert (pm.forward(obs[0, :8].contiguous()).cpu() - mean_ref).abs().max() <= 3e-5
    tr.close(); tr2.close(); pm.close()


def test_lmpc_plant_step_matches_oracle_model(built):
    """dart_lmpc_plant_step = one RK4 step of the reference's 8-state ODE (oracle/models.lmpc_step, rlmpc2.py:260-436)."""
    import torch
    from oracle import models
    rng = np.random.default_rng(9)
    B = 257
    x = rng.uniform(-0.1, 0.1, (B, 8))
    u = rng.uniform(-0.4, 0.4, (B, 2))
    pvec = np.clip(1.0 + 0.3 * rng.standard_normal((B, 34)), 0.01, 1.9)
    aux = np.concatenate([np.zeros((B, 2)), pvec], axis=1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = dart_b200.lmpc_plant_step(t(x), t(u), t(aux)).cpu().numpy()
    ref = np.stack([np.asarray(models.lmpc_step(x[b], u[b], pvec[b], 0.002)).reshape(-1) for b in range(B)])
    assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


def test_closed_loop_training_runs_and_matches_step_semantics(built):
    """LMPCTrainer on the surrogate plant: transitions are recorded every 8th step, a rollout of T transitions triggers
    epochs x minibatches optimiser steps, actions are the learner's samples, rewards equal the oracle's on the logged states."""
    import torch
    B, T = 64, 4
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    rng = np.random.default_rng(1)
    true_aux = torch.from_numpy(np.concatenate([np.zeros((B, 2)), np.clip(c["pvec"] + 0.2 * rng.standard_normal((B, 34)), 0.05, 1.8)], axis=1)).cuda()
    ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=3)
    ppo = dart_b200.PPOTrainer(capacity=256, epochs=2, mini_batch_size=128, reward_cfg=dict(max_delta=0.02, w_pos=40.0))
    g = torch.Generator(device="cuda").manual_seed(7)
    tr = dart_b200.LMPCTrainer(ctl, ppo, rollout_len=T, record_every=8, generator=g)
    x = torch.from_numpy(c["state"]).cuda(); tg = torch.from_numpy(c["target"]).cuda()
    p0 = ppo.state_dict()
    steps = 8 * T
    for k in range(steps):
        ctrl_before, prev_before = ctl.u_prev.cpu().numpy().copy(), tr.prev_cmd.cpu().numpy().copy()
        es, tp = tr.episode_step.cpu().numpy().copy(), tr.time_penalty.cpu().numpy().copy()
        u0, rew, done = tr.step(x, tg)
        a = tr._last[1].cpu().numpy()
        for b in (0, 31):
            rr, dd, _, _ = oppo.reward(x[b].cpu().numpy(), tg[b].cpu().numpy(), ctrl_before[b], prev_before[b], a[b], 1.0,
                                       int(es[b]), float(tp[b]), max_delta=0.02, w_pos=40.0)
            assert abs(float(rew[b]) - rr) <= 1e-5 * max(1.0, abs(rr)) and bool(done[b]) == dd
        assert torch.equal(ctl.action, tr._last[1])                 # the sampled action is what updated the model parameters
        assert tr.k == ((k // 8 + 1) % T if k % 8 == 0 else tr.k)
        x = dart_b200.lmpc_plant_step(x, u0, true_aux)
    assert tr.updates == 2 * (T * B // 128) and len(tr.mean_reward) == 1 and tr.k == 0
    p1 = ppo.state_dict()
    assert all(np.abs(p1[k] - p0[k]).max() > 0 for k in p0)         # every tensor was trained
    assert np.isfinite(x.cpu().numpy()).all() and (ctl.status.cpu().numpy() != dart_b200.STATUS_NUMERIC).all()
    ctl.engine.close(); ctl.policy.close(); ppo.close()


@pytest.mark.parametrize("M", [37, 1000])
def test_fused_tile_kernel_matches_unfused_chain(built, M, monkeypatch):
    """ppo_mid_kernel (layers 2-3 + loss + gradients per 64-sample tile) against the first, unfused GEMM chain
    (DART_PPO_UNFUSED=1): same arithmetic per element, different summation grouping over the samples."""
    pol = _perturbed_policy()
    obs, eps, action, logp, value, mean, old_logp, adv, ret = _rollout(M, M + 3, pol)
    d = [t.cuda() for t in (obs, action, old_logp, adv, ret)]
    grads, stats = [], []
    for unfused in (False, True):
        if unfused:
            monkeypatch.setenv("DART_PPO_UNFUSED", "1")
        else:
            monkeypatch.delenv("DART_PPO_UNFUSED", raising=False)
        tr = dart_b200.PPOTrainer(capacity=M, state_dict=pol.state_dict())
        stats.append(tr.update_minibatch(*d, apply=False).cpu().numpy().copy())
        grads.append(tr.gradient())
        tr.close()
    assert np.abs(stats[0][:3] - stats[1][:3]).max() <= 1e-5 * max(1.0, np.abs(stats[1][:3]).max())
    for k in grads[0]:
        ref = grads[1][k]
        assert np.abs(grads[0][k] - ref).max() <= 1e-7 + 1e-4 * np.abs(ref).max(), k      # incl. the split-K layer-1 forward of the fused path


def test_rlmpc_facade_training_mode(built, tmp_path):
    """RLMPC(model, data, {"train": True, ...}) as run.py builds it: sampled actions, rollout of `rollout_len` transitions (one per 8
    control steps), PPO update, checkpoints in the reference's format when an episode ends, reset request to the caller."""
    import torch
    model = dart_b200.GravityModel(-9.81, 0.002); data = dart_b200.StateHolder()
    b = data.body("cube2"); b.xmat = np.eye(3).reshape(-1); b.xpos[:] = [0.01, -0.02, 0.43]
    ckdir = str(tmp_path / "ckpt")
    params = {"Ts": 0.002, "nx": 8, "nu": 2, "N": 20, "Q": [200.0, 2.0, 200.0, 2.0, 0, 0, 0, 0], "Qt": [200.0, 2.0, 200.0, 2.0, 0, 0, 0, 0],
              "R": [0.1, 0.1, 1.0, 1.0], "u_bounds": (-0.4, 0.4), "body_name": "cube2", "g": 9.81, "max_param_abs": 2.0,
              "max_delta_abs": 0.02, "train": True, "seed": 0, "checkpoint_dir": ckdir, "rollout_len": 3, "epochs": 2,
              "mini_batch_size": 2, "max_episode_steps": 30, "w_pos": 40.0}
    with dart_b200.RLMPC(model, data, params) as ctl:
        assert ctl.training and ctl._trainer is not None
        p0 = ctl._ppo.state_dict()
        tgt = np.array([0.05, 0, 0.05, 0, 0, 0, 0, 0])
        k0 = ctl.views["model_params"].copy()
        for step in range(30):
            u, loss = ctl.solve(tgt)
            assert u.shape == (2,) and np.all(np.abs(u) <= 0.4 + 1e-12)
            if step == 0:
                assert not np.array_equal(ctl.views["model_params"], k0)       # the sampled action moved the parameters
        assert ctl._trainer.updates == 2 * 2                        # one rollout of 3 transitions -> 2 epochs x 2 minibatches (2 + 1)
        assert ctl.events["reset"].is_set() and ctl.episode_count == 1   # max_episode_steps reached at the 30th step
        p1 = ctl._ppo.state_dict()
        assert any(np.abs(p1[k] - p0[k]).max() > 0 for k in p0)
    for name in ("best_agent.pth", "latest_agent.pth"):
        ck = torch.load(os.path.join(ckdir, name), map_location="cpu", weights_only=True)
        pol = oppo.Policy()
        pol.load_state_dict(ck["model"])
        assert ck["episode"] == 1 and all(np.array_equal(ck["model"][k].numpy(), p1[k]) for k in p1)
    # evaluation run picks the trained actor up from the checkpoint directory, as the reference does (rlmpc2.py:567-573)
    params_eval = dict(params, train=False)
    with dart_b200.RLMPC(model, data, params_eval) as ev:
        assert not ev.training
        W = ev._batch.policy.weights[0][0]
        assert np.array_equal(W, p1["mean_net.0.weight"])


def test_graph_replay_is_bitwise_equal_to_eager(built):
    """train_rollout(graph=True): the minibatch step replayed from a CUDA graph (device-resident Adam step count, indices from
    a static buffer) gives the same parameters, moments and step count as the eager launches, bit for bit; capture itself
    leaves the learner untouched."""
    import torch
    T, B = 8, 40                                             # 320 transitions, minibatch 64 -> 5 full steps per epoch
    g0 = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn(T, B, 520, device="cuda", generator=g0)
    eps = torch.randn(T, B, 34, device="cuda", generator=g0)
    rew = torch.randn(T, B, device="cuda", generator=g0)
    done = (torch.rand(T, B, device="cuda", generator=g0) < 0.05).float()
    outs = []
    for graph in (False, True, True):
        tr = dart_b200.PPOTrainer(capacity=64, epochs=3, mini_batch_size=64)
        a, lp, v, _ = zip(*[tr.act(obs[t], eps[t]) for t in range(T)])
        act, logp, val = torch.stack(a), torch.stack(lp), torch.stack(v)
        last = val[-1].clone()
        steps = 0
        for rollout in range(2):                            # the second rollout reuses the captured graph
            gen = torch.Generator(device="cuda").manual_seed(5 + rollout)
            steps += tr.train_rollout(obs, act, logp, rew, val, done, last, generator=gen, graph=graph)
        assert steps == 2 * 3 * 5
        assert tr.graph_replays == (steps if graph else 0)
        p, m, vv, st = tr._get()
        assert st == steps
        outs.append((p, m, vv))
        tr.close()
    for k in range(3):
        assert np.array_equal(outs[0][k], outs[1][k]) and np.array_equal(outs[1][k], outs[2][k])


def test_reference_trained_policy_golden(built):
    """The policy the reference itself trained (tests/golden/ppo_reference_checkpoint.npz = best_agent.pth['model'], generated by
    tests/golden/make_ppo_golden.py) through the device path: rollout-time outputs, loss terms and the minibatch gradient against
    the committed torch results."""
    import torch
    from tests.helpers import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "ppo_reference_checkpoint.npz"))
    sd = dart_b200.unpack_params(g["params"])
    M = g["mean"].shape[0]
    gen = torch.Generator().manual_seed(2024)
    obs = torch.randn(M, 520, generator=gen); eps = torch.randn(M, 34, generator=gen)
    tr = dart_b200.PPOTrainer(capacity=M, state_dict=sd)
    a, lp, v, mu = tr.act(obs.cuda(), eps.cuda())
    for name, got in (("mean", mu), ("value", v), ("logp", lp), ("action", a)):
        ref = g[name]
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), name
    t = lambda k: torch.from_numpy(g[k]).cuda()
    stats = tr.update_minibatch(obs.cuda(), t("action"), t("old_logp"), t("adv"), t("ret"), apply=False).cpu().numpy()
    assert np.abs(stats[:3] - g["stats"][:3]).max() <= 1e-4 * max(1.0, np.abs(g["stats"][:3]).max())
    grad = tr.gradient()
    for k in grad:
        gf = grad[k].reshape(-1)
        n_ref = float(g["gnorm/" + k])
        assert abs(np.linalg.norm(gf.astype(np.float64)) - n_ref) <= 1e-4 * n_ref + 1e-9, k
        assert np.abs(gf[g["gidx/" + k]] - g["gval/" + k]).max() <= 1e-6 + 2e-4 * np.abs(grad[k]).max(), k
    tr.close()


def test_data_parallel_step_two_gpus(built):
    """One process per GPU, gradients all-reduced over NCCL (tests/dist_ppo_check.py): replicas bitwise identical, the step equals the
    single-GPU step on the concatenated minibatch."""
    import subprocess
    import sys
    import torch
    from tests.helpers import ROOT
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", os.path.join(ROOT, "tests", "dist_ppo_check.py")], capture_output=True, text=True, timeout=300)
    assert "DIST_PPO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_training_example_runs(built, tmp_path):
    """examples/lmpc_train_surrogate.py end to end (small): rollouts complete, a checkpoint in the reference's format is written."""
    import subprocess
    import sys
    import torch
    from tests.helpers import ROOT
    ck = str(tmp_path / "agent.pth")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "lmpc_train_surrogate.py"), "--instances", "64", "--rollouts", "2",
                        "--rollout-len", "4", "--mini-batch-size", "128", "--epochs", "2", "--checkpoint", ck],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "rollout   2" in r.stdout and "checkpoint written" in r.stdout
    sd = torch.load(ck, map_location="cpu", weights_only=True)["model"]
    oppo.Policy().load_state_dict(sd)
