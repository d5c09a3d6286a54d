"""LMPC policy kernels (tcgen05 MLP, observation build, parameter update) and the device-resident LMPC step."""
import numpy as np
import pytest

import dart_b200
from oracle import ipm, policy, problems

pytestmark = pytest.mark.gpu

# Default arithmetic ("fp32"): layer 1 as a 3xTF32 tensor-core product (A_hi W_hi + A_hi W_lo + A_lo W_hi, FP32
# accumulate), layers 2-3 in FP32 FMAs, tanh from the SFU exponential (<= 3e-7).  Measured against an fp64 reference:
# <= 1.0e-5 on O(1) outputs (torch's own fp32 forward: 2e-6).  Bound: 2e-5 absolute.
TOL_MLP = 2e-5
# The optional single-pass TF32 kernel (precision="tf32") multiplies TF32-truncated operands: observed 3-5e-3, bound 8e-3.
TOL_MLP_TF32 = 8e-3


def _torch_ref(weights, obs, dtype):
    import torch
    h = torch.from_numpy(obs).to(dtype)
    for i, (W, b) in enumerate(weights):
        h = h @ torch.from_numpy(W).to(dtype).T + torch.from_numpy(b).to(dtype)
        if i < 2:
            h = torch.tanh(h)
    return h.numpy()


@pytest.mark.parametrize("B", [1, 127, 128, 1000, 16384])
def test_policy_mlp_vs_torch(built, B):
    import torch
    rng = np.random.default_rng(B)
    weights = dart_b200.init_policy_weights(seed=3)
    weights = [(W, (0.05 * rng.standard_normal(b.shape)).astype(np.float32)) for W, b in weights]   # exercise the bias path
    obs = rng.standard_normal((B, 520)).astype(np.float32)
    pol = dart_b200.PolicyMLP(weights, device=0)
    out = pol.forward(torch.from_numpy(obs).cuda()).cpu().numpy()
    ref32 = _torch_ref(weights, obs, torch.float32)
    ref64 = _torch_ref(weights, obs, torch.float64)
    assert np.abs(ref32 - policy.mlp_forward(obs, weights)).max() < 1e-4          # oracle restatement == torch
    err = np.abs(out - ref64).max()
    print(f"B={B}: max|mean - fp64 ref| = {err:.2e}  (fp32 torch vs fp64: {np.abs(ref32 - ref64).max():.2e})")
    assert err <= TOL_MLP
    fast = dart_b200.PolicyMLP(weights, device=0, precision="tf32")
    err_fast = np.abs(fast.forward(torch.from_numpy(obs).cuda()).cpu().numpy() - ref64).max()
    assert TOL_MLP < err_fast <= TOL_MLP_TF32 or B == 1          # the single-pass kernel really is the coarser one


def test_policy_mlp_with_reference_checkpoint_shapes(built):
    """Real-architecture weights (random here; the reference's .pth files are untrusted pickles we do not ship)."""
    import torch
    weights = policy.orthogonal_policy_weights(seed=3)
    mine = dart_b200.init_policy_weights(seed=3)
    for (a, b), (c, d) in zip(weights, mine):
        assert np.array_equal(a, c) and np.array_equal(b, d)


def test_obs_push_and_param_update_vs_oracle(built):
    import torch
    import ctypes as C
    rng = np.random.default_rng(5)
    B = 33
    dev = torch.device("cuda", 0)
    L = dart_b200._lib.lib()
    norm = policy.ObsNormalizer(B)
    mean = torch.zeros((B, 52), dtype=torch.float64, device=dev); M2 = torch.zeros_like(mean)
    obs = [torch.zeros((B, 520), dtype=torch.float32, device=dev) for _ in range(2)]
    p = lambda t: C.c_void_p(t.data_ptr())
    for step in range(1, 14):
        st, tg = rng.standard_normal((B, 8)), rng.standard_normal((B, 8))
        ct, k = 0.3 * rng.standard_normal((B, 2)), rng.uniform(0.05, 1.8, (B, 34))
        ref = norm.push(st, tg, ct, k)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d_st, d_tg, d_ct, d_k = t(st), t(tg), t(ct), t(k)          # keep the device buffers alive across the launch
        rc = L.dart_policy_obs_push(B, step, p(d_st), p(d_tg), p(d_ct), p(d_k), 34, p(mean), p(M2), p(obs[0]), p(obs[1]), None)
        assert rc == 0
        torch.cuda.synchronize()
        obs = [obs[1], obs[0]]
        got = obs[0].cpu().numpy()
        assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), step
    # parameter update
    act = (0.5 * rng.standard_normal((B, 34))).astype(np.float32)
    k = rng.uniform(0.02, 1.85, (B, 34))
    kd = torch.from_numpy(k.copy()).to(dev)
    d_act = torch.from_numpy(act).to(dev)
    assert L.dart_policy_param_update(B, p(d_act), p(kd), 34, 2.0, 0.02, 1e-2, 0.1, 0.5, None) == 0
    ref = policy.write_params(policy.param_update(k, act, 2.0, 0.02, 1e-2), k, 2.0, 1e-2, 0.1, 0.5)
    assert np.abs(kd.cpu().numpy() - ref).max() <= 1e-6


def test_lmpc_batch_step_matches_oracle_pipeline(built):
    """Three steps of LMPCBatch: pvec after the update and the solve's u0/J against the oracle run on the same inputs."""
    import torch
    B = 48
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    dev = torch.device("cuda", 0)
    weights = dart_b200.init_policy_weights(seed=3)
    ctl = dart_b200.LMPCBatch(B, c["pvec"], weights=weights, device=0)
    norm = policy.ObsNormalizer(B)
    k = c["pvec"].copy(); control = np.zeros((B, 2)); Xw = Uw = None
    state = c["state"].copy()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for step in range(3):
        u_gpu = ctl.step(t(state), t(c["target"])).cpu().numpy()
        obs = norm.push(state, c["target"], control, c["pvec"])      # current_k of the observation is the stale start-up copy (rlmpc2.py:649)
        a = policy.mlp_forward(obs, weights)
        if step % 8 == 0:
            k = policy.write_params(policy.param_update(k, a, 2.0, 0.02, 1e-2), k, 2.0, 1e-2, 0.1, 0.5)
        assert np.abs(ctl.pvec.cpu().numpy() - k).max() <= 1e-6       # FP32-fidelity action -> 0.02 * 2e-5 in logit space
        prob = problems.lmpc_problem(state, control, ctl.pvec.cpu().numpy(), c["target"])   # same pvec: isolates the solve
        X0 = np.zeros((B, 21, 10)) if Xw is None else Xw
        U0 = np.zeros((B, 20, 2)) if Uw is None else Uw
        sol = ipm.solve(prob, X0=X0, U0=U0)
        assert (sol["status"] == 0).all() and (ctl.status.cpu().numpy() == 0).all()
        assert np.abs(u_gpu - sol["U"][:, 0]).max() <= 1e-4
        assert (np.abs(ctl.J.cpu().numpy() - sol["J"]) / np.abs(sol["J"])).max() <= 1e-6
        Xw, Uw = sol["X"], sol["U"]
        control = sol["U"][:, 0].copy()
        state = state + 0.002 * np.concatenate([state[:, 1:2], 0 * state[:, :1], state[:, 3:4], 0 * state[:, :1], 0 * state[:, :4]], axis=1)


def test_rlmpc_facade(built):
    model = dart_b200.GravityModel(-9.81, 0.002); data = dart_b200.StateHolder()
    b = data.body("cube2"); b.xmat = np.eye(3).reshape(-1); b.xpos[:] = [0.01, -0.02, 0.43]
    params = {"Ts": 0.002, "nx": 8, "nu": 2, "N": 20, "Q": [200.0, 2.0, 200.0, 2.0, 0, 0, 0, 0], "Qt": [200.0, 2.0, 200.0, 2.0, 0, 0, 0, 0],
              "R": [0.1, 0.1, 1.0, 1.0], "u_bounds": (-0.4, 0.4), "body_name": "cube2", "g": 9.81, "max_param_abs": 2.0,
              "max_delta_abs": 0.02, "train": False, "seed": 0, "checkpoint_dir": "/nonexistent"}
    with dart_b200.RLMPC(model, data, params) as ctl:
        tgt = np.array([0.05, 0, 0.05, 0, 0, 0, 0, 0])
        u, loss = ctl.solve(tgt)
        assert u.shape == (2,) and np.all(np.abs(u) <= 0.4) and ctl.views["w_opt"].shape == (208,)
        assert np.array_equal(ctl.views["control"], u)
        u2, _ = ctl.solve(tgt)
        assert np.all(np.isfinite(u2))


def test_lmpc_pipeline_end_to_end_within_the_tilt_bar(built):
    """The TF32 policy feeds the NLP through the parameter update.  17 closed-loop steps (policy updates at steps 0, 8,
    16) on the GPU against a pipeline that is the oracle end to end (float32 MLP, its own parameters, its own solves):
    the tilt commands must still agree within the north-star bar of 1e-4 rad (measured: 2e-6).  The parameters are
    only loosely bounded (2e-3, measured 6e-4): the reference's Welford normaliser divides near-constant channels by a
    std floor of 1e-6, which amplifies 1e-9 differences between the two state trajectories far more than TF32 does."""
    import torch
    B, T = 24, 17
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    dev = torch.device("cuda", 0)
    weights = dart_b200.init_policy_weights(seed=3)
    ctl = dart_b200.LMPCBatch(B, c["pvec"], weights=weights, device=0)
    norm = policy.ObsNormalizer(B)
    k = c["pvec"].copy(); control = np.zeros((B, 2)); Xw = Uw = None
    x_gpu = c["state"].copy(); x_ora = c["state"].copy()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    worst_u = worst_k = 0.0
    for step in range(T):
        u_gpu = ctl.step(t(x_gpu), t(c["target"])).cpu().numpy()
        obs = norm.push(x_ora, c["target"], control, c["pvec"])
        a = policy.mlp_forward(obs, weights)
        if step % 8 == 0:
            k = policy.write_params(policy.param_update(k, a, 2.0, 0.02, 1e-2), k, 2.0, 1e-2, 0.1, 0.5)
        prob = problems.lmpc_problem(x_ora, control, k, c["target"])
        sol = ipm.solve(prob, X0=np.zeros((B, 21, 10)) if Xw is None else Xw, U0=np.zeros((B, 20, 2)) if Uw is None else Uw)
        assert (sol["status"] == 0).all() and (ctl.status.cpu().numpy() == 0).all()
        Xw, Uw = sol["X"], sol["U"]
        control = sol["U"][:, 0].copy()
        worst_u = max(worst_u, np.abs(u_gpu - control).max())
        worst_k = max(worst_k, np.abs(ctl.pvec.cpu().numpy() - k).max())
        x_ora = sol["X"][:, 1, :8].copy()                      # plant = model prediction, each pipeline its own
        x_gpu = ctl.w.cpu().numpy()[:, 8:16].copy()
    print(f"end-to-end over {T} steps: max|du0| = {worst_u:.2e} rad, max|dpvec| = {worst_k:.2e}")
    assert worst_u <= 1e-4 and worst_k <= 2e-3
