"""Committed golden vectors (tests/golden, produced by make_golden.py from the oracle at tol 1e-10)."""
import os

import numpy as np
import pytest

import dart_b200
from oracle import ipm, policy, problems, rls
from tests import helpers

G = os.path.join(helpers.ROOT, "tests", "golden")
load = lambda n: np.load(os.path.join(G, n))


def _close(out, g, what):
    assert (out["status"] == 0).all(), what
    assert np.abs(out["u0"] - g["u0"]).max() <= helpers.TOL_U0, what
    assert (np.abs(out["J"] - g["J"]) / np.maximum(np.abs(g["J"]), 1e-9)).max() <= helpers.TOL_J, what


def test_oracle_reproduces_golden():
    g = load("pmpc_config2_s4.npz")
    a = g["aux"]
    r = ipm.solve(problems.pmpc_problem(g["state"], g["target"], Qp=a[:, 0], Qv=a[:, 1], R=a[:, 2], mu=a[:, 3]))
    _close(dict(u0=r["U"][:, 0], J=r["J"], status=r["status"]), g, "pmpc")
    g = load("rls_traj.npz")
    th = np.zeros((8, 2, 7)); P = np.tile(np.eye(7) * 1e3, (8, 2, 1, 1))
    for t in range(g["phi"].shape[0]):
        th, P = rls.rls_update_batch(th, P, g["phi"][t], g["y"][t], 0.995)
    assert np.array_equal(th, g["theta"][-1]) and np.array_equal(P, g["P_final"])
    g = load("policy_mlp.npz")
    w = [(g[f"W{i}"], g[f"b{i}"]) for i in range(3)]
    assert np.abs(policy.mlp_forward(g["obs"], w, np.float32) - g["mean"]).max() < 1e-5


@pytest.mark.parametrize("name,cfgf", [("pmpc_config2_s4.npz", "pmpc_cfg"), ("rmpc_b32.npz", "rmpc_cfg"), ("lmpc_b32.npz", "lmpc_cfg")])
def test_hostemu_vs_golden(hostemu, name, cfgf):
    g = load(name)
    x0 = g["state"] if "state" in g else g["x0"]
    ref = g["target"] if "target" in g else g["ref"]
    _close(hostemu.solve(getattr(dart_b200, cfgf)(), x0, ref, g["aux"]), g, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name,cfgf", [("pmpc_config2_s4.npz", "pmpc_cfg"), ("pmpc_config1.npz", "pmpc_cfg"),
                                       ("rmpc_b32.npz", "rmpc_cfg"), ("lmpc_b32.npz", "lmpc_cfg")])
def test_gpu_vs_golden(built, name, cfgf):
    g = load(name)
    x0 = g["state"] if "state" in g else g["x0"]
    ref = g["target"] if "target" in g else g["ref"]
    eng = dart_b200.NMPCEngine(getattr(dart_b200, cfgf)(), device=0)
    _close(eng.solve(x0, ref, aux=g["aux"] if "aux" in g else None, want_w=False), g, name)


@pytest.mark.gpu
def test_gpu_rls_and_mlp_vs_golden(built):
    import torch
    dev = torch.device("cuda", 0)
    g = load("rls_traj.npz")
    th = torch.zeros((8, 2, 7), dtype=torch.float64, device=dev)
    P = (torch.eye(7, dtype=torch.float64, device=dev) * 1e3).repeat(8, 2, 1, 1).contiguous()
    worst = 0.0
    for t in range(g["phi"].shape[0]):
        dart_b200.rls_update_device(th, P, torch.from_numpy(g["phi"][t]).to(dev), torch.from_numpy(g["y"][t]).to(dev), 0.995)
        if t % 40 == 39:
            worst = max(worst, np.abs(th.cpu().numpy() - g["theta"][t]).max() / np.abs(g["theta"][t]).max())
    assert worst <= helpers.TOL_RLS
    g = load("policy_mlp.npz")
    pol = dart_b200.PolicyMLP([(g[f"W{i}"], g[f"b{i}"]) for i in range(3)], device=0)
    out = pol.forward(torch.from_numpy(g["obs"]).to(dev)).cpu().numpy()
    assert np.abs(out - g["mean"]).max() <= 2e-5      # FP32-fidelity forward, see tests/test_gpu_lmpc_policy.py
