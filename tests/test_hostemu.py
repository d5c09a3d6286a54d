"""Solver logic of the CUDA source, compiled for the host with a 1-lane tile, against the oracle.

This is how the interior-point/Riccati code is unit-tested without a GPU; the GPU parity tests proper
(test_gpu_*.py) go through the C ABI.
"""
import numpy as np
import pytest

import dart_b200
from oracle import ipm
from tests import helpers


@pytest.fixture(params=["monotone", "mehrotra"])
def strategy(request, monkeypatch):
    """Barrier strategy of the host build (opts.h reads DART_BARRIER_MONOTONE / DART_BARRIER_MEHROTRA per call; the default,
    DART_BARRIER_AUTO, is the monotone schedule for RMPC) and the oracle's matching option:
    RMPC is one undivided NLP on both sides, so the two run the same iterates under either strategy."""
    monkeypatch.delenv("DART_BARRIER_MONOTONE", raising=False)
    monkeypatch.delenv("DART_BARRIER_MEHROTRA", raising=False)
    monkeypatch.setenv("DART_BARRIER_MONOTONE" if request.param == "monotone" else "DART_BARRIER_MEHROTRA", "1")
    return 1 if request.param == "mehrotra" else 0


def test_pmpc_matches_oracle(hostemu):
    c, aux, p = helpers.pmpc_case(4)
    out = hostemu.solve(dart_b200.pmpc_cfg(), c["state"], c["target"], aux)
    helpers.assert_parity(out, ipm.solve(p), "pmpc")


def test_pmpc_cfg_defaults_when_no_aux(hostemu):
    c1 = dart_b200.workloads.pmpc_config1()
    out = hostemu.solve(dart_b200.pmpc_cfg(), c1["state"], c1["target"], None)
    from oracle import problems
    ref = ipm.solve(problems.pmpc_problem(c1["state"], c1["target"], Qp=400, Qv=2, R=0.2, mu=0.1))
    helpers.assert_parity(out, ref, "pmpc config 1")


def test_rmpc_matches_oracle_iterate_for_iterate(hostemu, strategy):
    d, p = helpers.rmpc_case(32)
    out = hostemu.solve(dart_b200.rmpc_cfg(), d["x0"], d["ref"], d["aux"])
    ref = ipm.solve(p, opts=ipm.Options(mehrotra=strategy))
    helpers.assert_parity(out, ref, "rmpc")
    # same formulation, same algorithm: iteration counts coincide
    assert np.abs(out["iters"] - ref["iters"]).max() <= 1
    if strategy:       # predictor-corrector steps: about a quarter fewer iterations than the monotone schedule
        assert out["iters"].mean() < 0.8 * ipm.solve(p)["iters"].mean()


def test_lmpc_predictor_corrector_matches_oracle_variant(hostemu):
    """LMPC splits into two axis problems here and is one coupled NLP in the oracle: same KKT point, and the same number
    of predictor-corrector iterations to within one."""
    d, p = helpers.lmpc_case(32)
    out = hostemu.solve(dart_b200.lmpc_cfg(), d["x0"], d["ref"], d["aux"])
    ref = ipm.solve(p, opts=ipm.Options(mehrotra=1))
    helpers.assert_parity(out, ref, "lmpc predictor-corrector")
    assert np.abs(out["iters"] - ref["iters"]).max() <= 1
    assert out["iters"].mean() < 0.8 * ipm.solve(p)["iters"].mean()


def test_lmpc_matches_oracle(hostemu):
    d, p = helpers.lmpc_case(32)
    out = hostemu.solve(dart_b200.lmpc_cfg(), d["x0"], d["ref"], d["aux"])
    helpers.assert_parity(out, ipm.solve(p), "lmpc")


def test_warm_start_layout_roundtrip(hostemu):
    d, p = helpers.rmpc_case(8)
    cfg = dart_b200.rmpc_cfg()
    a = hostemu.solve(cfg, d["x0"], d["ref"], d["aux"])
    b = hostemu.solve(cfg, d["x0"], d["ref"], d["aux"], warm=a["w"])
    assert (b["status"] == 0).all()
    assert np.abs(a["u0"] - b["u0"]).max() < 1e-5
    assert np.abs(a["J"] - b["J"]).max() <= 1e-7 * np.abs(a["J"]).max()


def test_rmpc_infeasible_x0_flagged(hostemu):
    d, _ = helpers.rmpc_case(4)
    x0 = d["x0"].copy(); x0[0, 1] = 0.3      # |v0| > vmax acts on the fixed x_0 (np_mpc...:124-127)
    out = hostemu.solve(dart_b200.rmpc_cfg(), x0, d["ref"], d["aux"])
    assert out["status"][0] == dart_b200.STATUS_INFEASIBLE and (out["status"][1:] == 0).all()


def test_max_iter_status(hostemu):
    c, aux, _ = helpers.pmpc_case(1)
    out = hostemu.solve(dart_b200.pmpc_cfg(max_iter=2), c["state"], c["target"], aux)
    assert (out["status"] == dart_b200.STATUS_MAXITER).all() and (out["iters"] == 2).all()


def test_acceptable_level_termination_matches_oracle(hostemu, strategy):
    """IPOPT's acceptable_tol / acceptable_iter (the reference sets them for LMPC, rlmpc2.py:486-488).  RMPC is one
    undivided NLP on both sides, so solver and oracle stop at the same iterate."""
    d, p = helpers.rmpc_case(32)
    tight = hostemu.solve(dart_b200.rmpc_cfg(), d["x0"], d["ref"], d["aux"])
    opts = dict(tol=1e-12, acceptable_tol=1e-3, acceptable_iter=3)        # tol out of reach: only the acceptable exit
    out = hostemu.solve(dart_b200.rmpc_cfg(**opts), d["x0"], d["ref"], d["aux"])
    ref = ipm.solve(p, opts=ipm.Options(tol=1e-12, acc_tol=1e-3, acc_iter=3, mehrotra=strategy))
    assert (out["status"] == dart_b200.STATUS_ACCEPTABLE).all() and (ref["status"] == ipm.STATUS_ACCEPTABLE).all()
    assert np.array_equal(out["iters"], ref["iters"])
    assert np.abs(out["u0"] - ref["U"][:, 0]).max() < 1e-9
    if not strategy:      # (predictor-corrector steps converge so fast that three acceptable iterates in a row can come later)
        assert (out["iters"] < tight["iters"]).all()
    assert np.abs(out["u0"] - tight["u0"]).max() < 5e-3                   # an early exit, not a different optimum


def test_reference_lmpc_solver_options(hostemu):
    """tol 1e-4 / acceptable 1e-3 x 5 / max_iter 50: fewer iterations, same plan to within the early-exit tolerance."""
    from dart_b200.config import LMPC_REFERENCE_SOLVER_OPTIONS as ro
    d, p = helpers.lmpc_case(32)
    tight = hostemu.solve(dart_b200.lmpc_cfg(), d["x0"], d["ref"], d["aux"])
    out = hostemu.solve(dart_b200.lmpc_cfg(**ro), d["x0"], d["ref"], d["aux"])
    assert np.isin(out["status"], (dart_b200.STATUS_CONVERGED, dart_b200.STATUS_ACCEPTABLE)).all()
    assert out["iters"].mean() < tight["iters"].mean() - 1.0
    assert np.abs(out["u0"] - tight["u0"]).max() < 2e-3
    assert (np.abs(out["J"] - tight["J"]) / np.abs(tight["J"])).max() < 1e-4


def test_dual_warm_start_matches_oracle(hostemu, strategy):
    """dart_set_dual_state semantics on the host build of the solver: second solve of a slightly moved RMPC problem,
    started from the first one's plan, slacks and multipliers, against the oracle given the same state."""
    d, p = helpers.rmpc_case(16)
    cfg = dart_b200.rmpc_cfg()
    dual = np.zeros((16, hostemu.ndual(cfg)))
    a = hostemu.solve(cfg, d["x0"], d["ref"], d["aux"], dual=dual)
    ra = ipm.solve(p, opts=ipm.Options(mehrotra=strategy))
    assert (a["status"] == 0).all() and (dual[:, 0] == 1.0).all()
    # the next control cycle: state advanced by a little, previous command as u_prev
    x1 = d["x0"] + 0.002 * np.stack([d["x0"][:, 1], 0 * d["x0"][:, 1], d["x0"][:, 3], 0 * d["x0"][:, 3]], axis=1)
    aux1 = d["aux"].copy(); aux1[:, :2] = a["u0"]
    from oracle import problems
    p1 = problems.rmpc_problem(x1, a["u0"], d["theta"], d["ref"])
    warm = dict(lam=ra["lam"], s=ra["s"], zl=ra["zl"], zu=ra["zu"], valid=ra["status"] == 0)
    rb = ipm.solve(p1, X0=ra["X"], U0=ra["U"], opts=ipm.Options(mu0=1e-6, mehrotra=strategy), warm=warm)
    dual[3, 0] = 0.0; warm["valid"][3] = False            # one instance without a usable dual state: mu not below 1e-4
    rb = ipm.solve(p1, X0=ra["X"], U0=ra["U"], opts=ipm.Options(mu0=1e-6, mehrotra=strategy), warm=warm)
    b = hostemu.solve(dart_b200.rmpc_cfg(mu_init=1e-6), x1, d["ref"], aux1, warm=a["w"], dual=dual)
    plain = hostemu.solve(dart_b200.rmpc_cfg(mu_init=1e-4), x1, d["ref"], aux1, warm=a["w"])
    assert (b["status"] == 0).all() and (rb["status"] == 0).all()
    assert np.abs(b["iters"] - rb["iters"]).max() <= 1
    assert np.abs(b["u0"] - rb["U"][:, 0]).max() < 1e-8
    assert np.abs(b["u0"] - plain["u0"]).max() < 1e-6
    if not strategy:      # the adaptive barrier parameter does not depend on mu_init: nothing to gain from the dual state
        assert b["iters"].sum() < plain["iters"].sum()


def test_lmpc_without_tilt_rate_cost_runs_the_monotone_schedule(hostemu):
    """The tiled predictor-corrector step recovers its inverse pivots from the tilt-rate coupling (Solver::feedforward); with
    R_da = R_db = 0 that coupling vanishes and the launch falls back to the monotone schedule (opts.h launch_opts)."""
    from oracle import problems
    d, _ = helpers.lmpc_case(8)
    cfg = dart_b200.lmpc_cfg()
    cfg.Rl[2] = 0.0; cfg.Rl[3] = 0.0
    out = hostemu.solve(cfg, d["x0"], d["ref"], d["aux"])
    assert (out["status"] == 0).all() and np.isfinite(out["u0"]).all()
    base = hostemu.solve(dart_b200.lmpc_cfg(), d["x0"], d["ref"], d["aux"])
    assert out["iters"].mean() > base["iters"].mean()          # monotone: more iterations than the default's predictor-corrector steps
