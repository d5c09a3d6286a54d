"""acceptable-level termination, shifted-plan warm start and the plan-shift fallback, on the GPU."""
import numpy as np
import pytest

import dart_b200
from tests import helpers
from dart_b200.config import LMPC_REFERENCE_SOLVER_OPTIONS
from oracle import ipm

pytestmark = pytest.mark.gpu


def test_acceptable_exit_matches_oracle_rmpc(built):
    """RMPC is one undivided NLP on both sides: same iterates, so the acceptable exit lands on the same iteration."""
    d, p = helpers.rmpc_case(128)
    opts = dict(tol=1e-12, acceptable_tol=1e-3, acceptable_iter=3)
    out = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(**opts), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    ref = ipm.solve(p, opts=ipm.Options(tol=1e-12, acc_tol=1e-3, acc_iter=3))
    assert (out["status"] == dart_b200.STATUS_ACCEPTABLE).all() and (ref["status"] == ipm.STATUS_ACCEPTABLE).all()
    assert np.array_equal(out["iters"], ref["iters"])
    assert np.abs(out["u0"] - ref["U"][:, 0]).max() < 1e-9


def test_reference_lmpc_options_fewer_iterations_same_plan(built):
    d, p = helpers.lmpc_case(256)
    tight = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    out = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(**LMPC_REFERENCE_SOLVER_OPTIONS), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    assert np.isin(out["status"], (dart_b200.STATUS_CONVERGED, dart_b200.STATUS_ACCEPTABLE)).all()
    assert out["iters"].mean() < tight["iters"].mean() - 1.0
    assert np.abs(out["u0"] - tight["u0"]).max() < 2e-3            # early exit at KKT error <= 1e-3/1e-4, not another optimum
    assert (np.abs(out["J"] - tight["J"]) / np.abs(tight["J"])).max() < 1e-4


def test_axis_status_combination_ranks_acceptable_below_failures(built):
    """One axis converged / acceptable, the other out of iterations: the instance reports the worse of the two."""
    d, _ = helpers.lmpc_case(16)
    out = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(tol=1e-12, acceptable_tol=1e-3, acceptable_iter=2, max_iter=3), device=0).solve(
        d["x0"], d["ref"], aux=d["aux"])
    assert np.isin(out["status"], (dart_b200.STATUS_ACCEPTABLE, dart_b200.STATUS_MAXITER)).all()
    assert (out["status"] == dart_b200.STATUS_MAXITER).any()


def _closed_loop(B, T, **kw):
    import torch
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=3, device=0, **kw)
    x = torch.from_numpy(c["state"]).cuda(); tg = torch.from_numpy(c["target"]).cuda()
    iters, us = 0, []
    for _ in range(T):
        u = ctl.step(x, tg)
        us.append(u.cpu().numpy().copy())
        iters += int(ctl.iters.sum().item())
        x = ctl.w[:, 8:16].contiguous()
    return ctl, np.stack(us), iters / (B * T)


def test_shifted_warm_start_same_commands(built):
    """The warm start changes the path of the iterations, not the optimum they reach."""
    _, ua, ia = _closed_loop(64, 12)
    _, ub, ib = _closed_loop(64, 12, warm_start="shift")
    assert np.abs(ua - ub).max() < 1e-5
    print(f"mean iterations: unshifted {ia:.2f}, shifted {ib:.2f}")


def test_plan_shift_fallback(built):
    """With an iteration cap no solve can meet after step 0, commands walk along the last good plan (rlmpc2.py:1013-1018)."""
    import torch
    B = 8
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=3, device=0, plan_fallback=True)
    x = torch.from_numpy(c["state"]).cuda(); tg = torch.from_numpy(c["target"]).cuda()
    u0 = ctl.step(x, tg).cpu().numpy().copy()
    plan = ctl.w[:, 8 * 21:].view(B, 20, 2).cpu().numpy().copy()
    assert (ctl.status.cpu().numpy() == 0).all() and np.array_equal(u0, plan[:, 0])
    ctl.engine = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(max_iter=1), device=0)       # every further solve hits the cap
    far = tg.clone(); far[:, 0] += 0.05                                                # move the target so one step is not enough
    for k in (1, 2, 3):
        u = ctl.step(x, far).cpu().numpy()
        assert (ctl.status.cpu().numpy() == dart_b200.STATUS_MAXITER).all()
        assert np.array_equal(u, plan[:, k])
    assert int(ctl.n_fallback.item()) == 3 * B
    assert np.array_equal(ctl.w[:, 8 * 21:].view(B, 20, 2).cpu().numpy(), plan)        # the good plan stays the warm start


def test_pmpc_episode_warm_start_same_metrics(built):
    """Warm start + warm-started barrier change the iteration path only: same episode to the solver tolerance."""
    c, aux = dart_b200.workloads.pmpc_inputs(2)
    cold = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, device=0)
    warm = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, device=0, warm_start=True)
    a, b = cold.run(60), warm.run(60)
    assert a["not_converged_solves"] == 0 and b["not_converged_solves"] == 0
    assert np.abs(cold.state.cpu().numpy() - warm.state.cpu().numpy()).max() < 1e-6
    assert b["mean_iters"] < 0.75 * a["mean_iters"]
    print(f"mean iterations per solve: cold {a['mean_iters']:.2f}, warm {b['mean_iters']:.2f}")
