"""acceptable-level termination, shifted-plan warm start and the plan-shift fallback, on the GPU."""
import numpy as np
import pytest

import dart_b200
from tests import helpers
from dart_b200.config import LMPC_REFERENCE_SOLVER_OPTIONS
from oracle import ipm

pytestmark = pytest.mark.gpu


def test_acceptable_exit_matches_oracle_rmpc(built):
    """RMPC is one undivided NLP on both sides: same iterates, so the acceptable exit lands on the same iteration -- under
    the monotone schedule and under the predictor-corrector steps a cold-started call runs by default."""
    d, p = helpers.rmpc_case(128)
    opts = dict(tol=1e-12, acceptable_tol=1e-3, acceptable_iter=3)
    for strategy, meh in (("monotone", 0), ("auto", 1)):
        eng = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(**opts), device=0)
        eng.set_barrier_strategy(strategy)
        out = eng.solve(d["x0"], d["ref"], aux=d["aux"])
        ref = ipm.solve(p, opts=ipm.Options(tol=1e-12, acc_tol=1e-3, acc_iter=3, mehrotra=meh))
        assert (out["status"] == dart_b200.STATUS_ACCEPTABLE).all() and (ref["status"] == ipm.STATUS_ACCEPTABLE).all()
        assert np.array_equal(out["iters"], ref["iters"]), strategy
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


def _closed_loop(B, T, strategy=None, **kw):
    import torch
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=3, device=0, **kw)
    if strategy:
        ctl.engine.set_barrier_strategy(strategy)
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


def test_dual_warm_start_rmpc_loop_matches_oracle(built):
    """RMPCBatch(dual_warm=True) against the oracle doing the same dual warm start: same iterates, same commands."""
    import torch
    from oracle import problems, rls
    from tests.test_gpu_rmpc_loop import _plant_step
    B, T = 16, 8
    c = dart_b200.workloads.rmpc_config3(B, seed=2)
    x = c["x0"].copy(); x[:, [1, 3]] *= 0.5
    P0 = 1.0
    ctl = dart_b200.RMPCBatch(B, c["target"], x, device=0, rls_P0=P0, dual_warm=True)
    ref_ctl = dart_b200.RMPCBatch(B, c["target"], x, device=0, rls_P0=P0)
    rv0 = np.zeros((B, 4)); rv0[:, [0, 2]] = x[:, [0, 2]]
    ctl.set_virtual_reference(rv0); ref_ctl.set_virtual_reference(rv0)
    th = np.zeros((B, 2, 7)); P = np.tile(np.eye(7) * P0, (B, 2, 1, 1))
    r_v = rv0.copy(); prev = x.copy(); u_prev = np.zeros((B, 2)); sol = None
    x_o = x.copy(); x_r = x.copy()
    it_dual = it_ref = 0
    for t in range(T):
        u_gpu = ctl.step(torch.from_numpy(x).cuda()).cpu().numpy()
        u_ref = ref_ctl.step(torch.from_numpy(x_r).cuda()).cpu().numpy()
        phi = rls.regressor(prev, 0.1); y = rls.accel_measurement(x_o, prev, 0.002)
        th, P = rls.rls_update_batch(th, P, phi, y, 0.995)
        r_v = problems.reference_governor(r_v, c["target"])
        ref = problems.build_ref_traj(x_o, r_v, c["target"], 20, 4, 0.2)
        prob = problems.rmpc_problem(x_o, u_prev, th.reshape(B, 14), ref)
        if sol is None:
            new = ipm.solve(prob, X0=np.zeros((B, 21, 6)), U0=np.zeros((B, 20, 2)))
        else:
            warm = dict(lam=sol["lam"], s=sol["s"], zl=sol["zl"], zu=sol["zu"], valid=np.isin(sol["status"], (0, 4)))
            new = ipm.solve(prob, X0=sol["X"], U0=sol["U"], opts=ipm.Options(mu0=1e-6), warm=warm)
        sol = new
        assert (sol["status"] == 0).all() and (ctl.status.cpu().numpy() == 0).all()
        assert np.abs(u_gpu - sol["U"][:, 0]).max() <= helpers.TOL_U0, (t, np.abs(u_gpu - sol["U"][:, 0]).max())
        assert np.abs(ctl.iters.cpu().numpy() - sol["iters"]).max() <= 1, (t, ctl.iters.cpu().numpy(), sol["iters"])
        if t > 0:
            it_dual += int(ctl.iters.sum().item()); it_ref += int(ref_ctl.iters.sum().item())
        prev = x_o.copy(); u_prev = sol["U"][:, 0].copy()
        x_o = _plant_step(x_o, sol["U"][:, 0], c["mu_plant"], c["c_plant"])
        x = _plant_step(x, u_gpu, c["mu_plant"], c["c_plant"])
        x_r = _plant_step(x_r, u_ref, c["mu_plant"], c["c_plant"])
    print(f"iterations per warm solve: dual {it_dual / (B * (T - 1)):.2f}, primal + mu {it_ref / (B * (T - 1)):.2f}")
    assert np.abs(x - x_r).max() < 1e-6                  # same closed loop as without the dual warm start
    assert it_dual < it_ref


def test_dual_warm_start_lmpc_loop(built):
    """The dual state saves iterations of the MONOTONE schedule (its barrier can start at 1e-6); the predictor-corrector
    steps LMPC runs by default pick mu themselves and need neither -- same commands either way."""
    ca, ua, ia = _closed_loop(64, 12, strategy="monotone", warm_mu=1e-4)
    cb, ub, ib = _closed_loop(64, 12, strategy="monotone", dual_warm=True)
    assert (cb.status.cpu().numpy() == 0).all()
    assert np.abs(ua - ub).max() < 1e-5
    print(f"LMPC mean iterations (monotone): primal + mu {ia:.2f}, dual {ib:.2f}")
    assert ib < ia
    cc, uc, ic = _closed_loop(64, 12)                       # default: predictor-corrector, primal warm start
    cd, ud, idd = _closed_loop(64, 12, dual_warm=True)      # ... with the dual state
    assert (cc.status.cpu().numpy() == 0).all() and (cd.status.cpu().numpy() == 0).all()
    assert np.abs(uc - ua).max() < 1e-5 and np.abs(ud - ua).max() < 1e-5
    print(f"LMPC mean iterations (predictor-corrector): primal {ic:.2f}, dual {idd:.2f}")
    assert ic < ia
    # a row whose solve failed is not reused: invalidate by hand and solve again
    cb.dual[:, 0] = 0.0
    import torch
    c = dart_b200.workloads.lmpc_config4(64, seed=3)
    cb.step(cb.w[:, 8:16].contiguous(), torch.from_numpy(c["target"]).cuda())
    assert (cb.status.cpu().numpy() == 0).all()
