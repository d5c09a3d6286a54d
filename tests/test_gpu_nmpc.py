"""GPU parity tests proper: the CUDA path through the C ABI against the oracle on the same seeded inputs.

Bars (BASELINE.json north_star): |u0 - u0_oracle| <= 1e-4 rad, |J - J_oracle| / |J_oracle| <= 1e-6.
"""
import numpy as np
import pytest

import dart_b200
from oracle import ipm, models, problems
from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pmpc_engine(built):
    return dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=0)


def test_pmpc_config2_full_batch_parity(pmpc_engine):
    """BASELINE config 2 at full size: 18 objects x 64 states = 1152 instances."""
    c, aux, p = helpers.pmpc_case(64)
    out = pmpc_engine.solve(c["state"], c["target"], aux=aux)
    ref = ipm.solve(p)
    du0, dJ = helpers.assert_parity(out, ref, "pmpc config 2")
    print(f"pmpc config2: B={p.B} max|du0|={du0:.2e} max rel dJ={dJ:.2e} iters mean={out['iters'].mean():.2f} max={out['iters'].max()}")


def test_pmpc_config1_single_instance(pmpc_engine):
    c1 = dart_b200.workloads.pmpc_config1()
    out = pmpc_engine.solve(c1["state"], c1["target"])
    ref = ipm.solve(problems.pmpc_problem(c1["state"], c1["target"], Qp=400, Qv=2, R=0.2, mu=0.1))
    helpers.assert_parity(out, ref, "pmpc config 1")


def test_pmpc_decision_vector_layout_and_z_rows(pmpc_engine):
    """w = [vec(X); vec(U)] (mpc_3d.py:69,137): x/y columns are dynamically feasible, z columns follow pmpc_step."""
    c, aux, p = helpers.pmpc_case(2)
    out = pmpc_engine.solve(c["state"], c["target"], aux=aux)
    N = 15
    X = out["w"][:, :(N + 1) * 6].reshape(-1, N + 1, 6)
    U = out["w"][:, (N + 1) * 6:].reshape(-1, N, 2)
    assert np.array_equal(U[:, 0], out["u0"])
    assert np.abs(X[:, 0] - c["state"]).max() == 0
    for k in range(N):
        nxt = models.pmpc_step(X[:, k], U[:, k], -9.81, c["mu"], 0.002)
        assert np.abs(nxt - X[:, k + 1]).max() < 1e-8
    assert (np.abs(U) <= 0.6).all()


@pytest.mark.parametrize("lanes", [2, 4, 8, 16])
def test_pmpc_lane_widths_agree(built, lanes):
    """Same algorithm (the monotone barrier schedule), different mapping of the horizon to lanes: same iterates up to
    summation order.  (The 16-lane tiles run predictor-corrector steps by default; that pair is compared below.)"""
    c, aux, p = helpers.pmpc_case(8)
    ref = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=0)
    ref.set_barrier_strategy("monotone")
    base = ref.solve(c["state"], c["target"], aux=aux)
    eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(lanes=lanes), device=0)
    eng.set_barrier_strategy("monotone")
    out = eng.solve(c["state"], c["target"], aux=aux)
    assert eng.last_launch_config()["lanes"] == lanes
    assert (out["status"] == 0).all()
    assert np.abs(out["u0"] - base["u0"]).max() < 1e-6
    assert (np.abs(out["J"] - base["J"]) / np.abs(base["J"])).max() < 1e-8
    assert np.array_equal(out["iters"], base["iters"])


def test_pmpc_barrier_strategies_reach_the_same_kkt_points(built):
    """dart_set_barrier_strategy: Mehrotra predictor-corrector steps (default on the 16-lane scan path) against the monotone
    schedule on the headline batch -- a different iterate path to the same KKT points (both against the oracle's, which
    runs the monotone method), in fewer iterations."""
    c, aux, p = helpers.pmpc_case(64)
    ref = ipm.solve(p)
    outs = {}
    for strat in ("monotone", "mehrotra"):
        eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=0)
        eng.set_barrier_strategy(strat)
        outs[strat] = eng.solve(c["state"], c["target"], aux=aux)
        assert eng.last_launch_config()["lanes"] == 16
        helpers.assert_parity(outs[strat], ref, "pmpc " + strat)
    mono, pc = outs["monotone"], outs["mehrotra"]
    assert np.abs(pc["u0"] - mono["u0"]).max() <= helpers.TOL_U0
    assert (np.abs(pc["J"] - mono["J"]) / np.abs(mono["J"])).max() <= helpers.TOL_J
    assert pc["iters"].max() < mono["iters"].max() and pc["iters"].mean() < 0.75 * mono["iters"].mean()
    # the oracle's own predictor-corrector variant (same algorithm, coupled 6-state problem, dense KKT solves)
    ref_pc = ipm.solve(p, opts=ipm.Options(mehrotra=1, mu0=0.1))        # PmpcAxis::MU0_PC
    assert (ref_pc["status"] == 0).all()
    assert abs(float(pc["iters"].mean()) - float(ref_pc["iters"].mean())) < 0.5


def test_rmpc_parity(built):
    """RMPC is one undivided NLP on both sides: same iterates as the oracle under the monotone schedule (and under
    predictor-corrector steps, next test)."""
    d, p = helpers.rmpc_case(256)
    eng = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(), device=0)
    eng.set_barrier_strategy("monotone")
    out = eng.solve(d["x0"], d["ref"], aux=d["aux"])
    ref = ipm.solve(p)
    du0, dJ = helpers.assert_parity(out, ref, "rmpc")
    assert np.abs(out["iters"] - ref["iters"]).max() <= 1
    print(f"rmpc: max|du0|={du0:.2e} max rel dJ={dJ:.2e} iters mean={out['iters'].mean():.2f}")


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_rmpc_lane_widths_agree(built, lanes):
    d, p = helpers.rmpc_case(64)
    base = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    out = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(lanes=lanes), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    assert (out["status"] == 0).all()
    assert np.abs(out["u0"] - base["u0"]).max() < 1e-6


def test_rmpc_warm_start_and_infeasible_flag(built):
    d, p = helpers.rmpc_case(16)
    eng = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(), device=0)
    a = eng.solve(d["x0"], d["ref"], aux=d["aux"])
    b = eng.solve(d["x0"], d["ref"], aux=d["aux"], warm_w=a["w"])
    assert (b["status"] == 0).all() and np.abs(a["u0"] - b["u0"]).max() < 1e-5
    x0 = d["x0"].copy(); x0[3, 3] = -0.35
    c = eng.solve(x0, d["ref"], aux=d["aux"])
    assert c["status"][3] == dart_b200.STATUS_INFEASIBLE


def test_lmpc_parity(built):
    d, p = helpers.lmpc_case(256)
    eng = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(), device=0)
    out = eng.solve(d["x0"], d["ref"], aux=d["aux"])
    ref = ipm.solve(p)
    du0, dJ = helpers.assert_parity(out, ref, "lmpc")
    print(f"lmpc: max|du0|={du0:.2e} max rel dJ={dJ:.2e} iters mean={out['iters'].mean():.2f}")


@pytest.mark.parametrize("lanes", [4, 8, 16])
def test_lmpc_lane_widths_agree(built, lanes):
    """The element-per-lane Riccati rounds take 1..8 passes depending on the tile width, and with 4 lanes (< nx) one
    lane runs the forward sweep instead of the tile; all widths must agree."""
    d, p = helpers.lmpc_case(64)
    base = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    out = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(lanes=lanes), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    assert (out["status"] == 0).all()
    assert np.abs(out["u0"] - base["u0"]).max() < 1e-6
    assert (np.abs(out["J"] - base["J"]) / np.abs(base["J"])).max() < 1e-8


def test_rmpc_predictor_corrector_matches_oracle_variant(built):
    """Predictor-corrector steps on the tiled-sweep path (pc_rows + corrector_tile) -- what DART_BARRIER_AUTO runs for RMPC
    calls without a warm plan.  RMPC is one undivided NLP on both sides, so the kernel and the oracle's predictor-corrector
    variant take the same iterates; against the monotone schedule the KKT points coincide in about 30 % fewer iterations.
    A warm-started call under AUTO runs the monotone schedule."""
    d, p = helpers.rmpc_case(256)
    eng = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(), device=0)
    auto = eng.solve(d["x0"], d["ref"], aux=d["aux"])
    eng.set_barrier_strategy("monotone")
    mono = eng.solve(d["x0"], d["ref"], aux=d["aux"])
    eng.set_barrier_strategy("mehrotra")
    out = eng.solve(d["x0"], d["ref"], aux=d["aux"])
    ref = ipm.solve(p, opts=ipm.Options(mehrotra=1))
    du0, dJ = helpers.assert_parity(out, ref, "rmpc predictor-corrector")
    assert np.abs(out["iters"] - ref["iters"]).max() <= 1
    assert np.abs(out["u0"] - ref["U"][:, 0]).max() < 1e-8
    assert np.abs(out["u0"] - mono["u0"]).max() <= helpers.TOL_U0
    assert (np.abs(out["J"] - mono["J"]) / np.abs(mono["J"])).max() <= helpers.TOL_J
    assert out["iters"].mean() < 0.8 * mono["iters"].mean()
    assert np.array_equal(auto["iters"], out["iters"]) and np.array_equal(auto["u0"], out["u0"])      # cold start: AUTO = these steps
    # warm-started calls: AUTO = the monotone schedule
    warm_m = eng.solve(d["x0"], d["ref"], aux=d["aux"], warm_w=mono["w"])
    eng.set_barrier_strategy("monotone")
    warm_ref = eng.solve(d["x0"], d["ref"], aux=d["aux"], warm_w=mono["w"])
    eng.set_barrier_strategy("auto")
    warm_a = eng.solve(d["x0"], d["ref"], aux=d["aux"], warm_w=mono["w"])
    assert np.array_equal(warm_a["iters"], warm_ref["iters"]) and np.array_equal(warm_a["u0"], warm_ref["u0"])
    assert (warm_m["status"] == 0).all() and np.abs(warm_m["u0"] - warm_ref["u0"]).max() < 1e-5


@pytest.mark.parametrize("method,lanes", [("rmpc", 8), ("rmpc", 16), ("lmpc", 4), ("lmpc", 8), ("lmpc", 16)])
def test_predictor_corrector_lane_widths_agree(built, method, lanes):
    """Sub-warp tiles run the predictor-corrector step in lockstep with the other tiles of their warp (finished tiles as
    ghosts); 4-lane LMPC tiles (< nx) run the corrector's recursions in one lane."""
    case, cfg = (helpers.rmpc_case, dart_b200.rmpc_cfg) if method == "rmpc" else (helpers.lmpc_case, dart_b200.lmpc_cfg)
    d, p = case(64)
    outs = []
    for ln in (0, lanes):
        eng = dart_b200.NMPCEngine(cfg(lanes=ln), device=0)
        eng.set_barrier_strategy("mehrotra")
        outs.append(eng.solve(d["x0"], d["ref"], aux=d["aux"]))
        if ln:
            assert eng.last_launch_config()["lanes"] == lanes
    base, out = outs
    assert (out["status"] == 0).all() and (base["status"] == 0).all()
    assert np.abs(out["u0"] - base["u0"]).max() < 1e-6
    assert (np.abs(out["J"] - base["J"]) / np.abs(base["J"])).max() < 1e-8
    assert np.abs(out["iters"] - base["iters"]).max() <= 1


def test_lmpc_barrier_strategies_reach_the_same_kkt_points(built):
    """LMPC's default (DART_BARRIER_AUTO) is the predictor-corrector step: same KKT points as the monotone schedule and
    as the oracle (coupled 10-state problem, either variant), in about a third fewer iterations."""
    d, p = helpers.lmpc_case(256)
    outs = {}
    for strat in ("monotone", "mehrotra", "auto"):
        eng = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(), device=0)
        eng.set_barrier_strategy(strat)
        outs[strat] = eng.solve(d["x0"], d["ref"], aux=d["aux"])
    ref_pc = ipm.solve(p, opts=ipm.Options(mehrotra=1))
    helpers.assert_parity(outs["mehrotra"], ref_pc, "lmpc predictor-corrector")
    helpers.assert_parity(outs["monotone"], ipm.solve(p), "lmpc monotone")
    assert np.array_equal(outs["auto"]["iters"], outs["mehrotra"]["iters"]) and np.array_equal(outs["auto"]["u0"], outs["mehrotra"]["u0"])
    assert np.abs(outs["mehrotra"]["u0"] - outs["monotone"]["u0"]).max() <= helpers.TOL_U0
    # (two axis problems here, one coupled problem there: the counts agree to within one on 255 of 256 instances, two on the last)
    assert np.abs(outs["mehrotra"]["iters"] - ref_pc["iters"]).max() <= 2
    assert abs(float(outs["mehrotra"]["iters"].mean()) - float(ref_pc["iters"].mean())) < 0.3
    assert outs["mehrotra"]["iters"].mean() < 0.75 * outs["monotone"]["iters"].mean()


def test_solve_into_reused_result_arrays(pmpc_engine):
    """NMPCEngine.solve(..., out=make_out(B)): same results as a call that allocates, written into the caller's arrays."""
    c, aux, p = helpers.pmpc_case(4)
    fresh = pmpc_engine.solve(c["state"], c["target"], aux=aux, want_w=False)
    out = pmpc_engine.make_out(p.B, want_w=False)
    u0_arr = out["u0"]
    r = pmpc_engine.solve(c["state"], c["target"], aux=aux, want_w=False, out=out)
    assert r is out and out["u0"] is u0_arr and out["w"] is None
    for k in ("u0", "J", "status", "iters"):
        assert np.array_equal(out[k], fresh[k]), k
    perm = np.arange(p.B)[::-1].copy()
    pmpc_engine.solve(c["state"][perm], c["target"][perm], aux=aux[perm], want_w=False, out=out)      # overwritten by the next call
    assert np.array_equal(out["u0"], fresh["u0"][perm])
    with pytest.raises(ValueError):
        pmpc_engine.solve(c["state"][:3], c["target"][:3], aux=aux[:3], want_w=False, out=out)


def test_device_pointer_entry_and_quaternion(pmpc_engine):
    import torch
    c, aux, p = helpers.pmpc_case(4)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = pmpc_engine.solve_device(t(c["state"]), t(c["target"]), aux=t(aux))
    torch.cuda.synchronize()
    host = pmpc_engine.solve(c["state"], c["target"], aux=aux, want_w=False)
    assert np.array_equal(out["u0"].cpu().numpy(), host["u0"]) and np.array_equal(out["J"].cpu().numpy(), host["J"])
    q = dart_b200.tilt_to_quat_device(out["u0"]).cpu().numpy()
    assert np.abs(q - models.tilt_to_quat(host["u0"])).max() < 1e-15


def test_batch_permutation_invariance(pmpc_engine):
    c, aux, p = helpers.pmpc_case(8)
    perm = np.random.default_rng(0).permutation(p.B)
    a = pmpc_engine.solve(c["state"], c["target"], aux=aux, want_w=False)
    b = pmpc_engine.solve(c["state"][perm], c["target"][perm], aux=aux[perm], want_w=False)
    assert np.array_equal(a["u0"][perm], b["u0"]) and np.array_equal(a["J"][perm], b["J"])


def test_empty_and_ragged_batches(pmpc_engine):
    c, aux, p = helpers.pmpc_case(1)
    for B in (1, 3, 17):
        out = pmpc_engine.solve(c["state"][:B], c["target"][:B], aux=aux[:B], want_w=False)
        assert out["u0"].shape == (B, 2) and (out["status"] == 0).all()
    out = pmpc_engine.solve(np.zeros((0, 6)), np.zeros((0, 6)), aux=np.zeros((0, 4)), want_w=False)
    assert out["u0"].shape == (0, 2)


def test_result_rows_and_capacity_check(pmpc_engine):
    import torch
    c, aux, p = helpers.pmpc_case(2)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rows = torch.zeros((p.B, 4), dtype=torch.float64, device=dev)
    pmpc_engine.set_result_rows(rows)
    try:
        out = pmpc_engine.solve_device(t(c["state"]), t(c["target"]), aux=t(aux))
        torch.cuda.synchronize()
        r = rows.cpu().numpy()
        assert np.array_equal(r[:, :2], out["u0"].cpu().numpy()) and np.array_equal(r[:, 2], out["J"].cpu().numpy())
        assert np.array_equal(r[:, 3], out["status"].cpu().numpy().astype(np.float64))
        small = torch.zeros((4, 4), dtype=torch.float64, device=dev)
        pmpc_engine.set_result_rows(small)
        with pytest.raises(dart_b200.DartError):
            pmpc_engine.solve_device(t(c["state"]), t(c["target"]), aux=t(aux))
    finally:
        pmpc_engine.set_result_rows(None)


@pytest.mark.parametrize("method", ["pmpc", "rmpc", "lmpc"])
def test_bitwise_repeatable_under_load(built, method):
    """compute-sanitizer is closed on this GPU pool; a missing tile-level sync would show as run-to-run nondeterminism,
    so: the same large batch solved five times (different co-resident warps each time) must agree bit for bit, and a
    small batch embedded in the large one must equal the same instances solved alone."""
    if method == "pmpc":
        c, aux, _ = helpers.pmpc_case(512)
        eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(lanes=8), device=0)     # fixed width: the auto choice depends on B
        args = (c["state"], c["target"]); kw = dict(aux=aux)
    elif method == "rmpc":
        d, _ = helpers.rmpc_case(2048)
        eng = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(lanes=32), device=0)
        args = (d["x0"], d["ref"]); kw = dict(aux=d["aux"])
    else:
        d, _ = helpers.lmpc_case(4096)
        eng = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(lanes=16), device=0)
        args = (d["x0"], d["ref"]); kw = dict(aux=d["aux"])
    first = eng.solve(*args, want_w=True, **kw)
    for _ in range(4):
        again = eng.solve(*args, want_w=True, **kw)
        for k in ("u0", "J", "w", "status", "iters"):
            assert np.array_equal(first[k], again[k]), k
    sub = eng.solve(args[0][:37], args[1][:37], aux=kw["aux"][:37], want_w=True)
    for k in ("u0", "J", "w", "status", "iters"):
        assert np.array_equal(first[k][:37], sub[k]), k


@pytest.mark.parametrize("N", [5, 30])
def test_runtime_horizon_kernels(built, N):
    """Horizons other than the reference's 15/20 take the runtime-N instantiation; parity must hold there too."""
    c, aux, _ = helpers.pmpc_case(2)
    out = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(N=N), device=0).solve(c["state"], c["target"], aux=aux)
    ref = ipm.solve(problems.pmpc_problem(c["state"], c["target"], Qp=c["Qp"], Qv=c["Qv"], R=c["R"], mu=c["mu"], N=N))
    helpers.assert_parity(out, ref, f"pmpc N={N}")
    assert out["w"].shape == (36, (N + 1) * 6 + 2 * N)
    d = dart_b200.workloads.rmpc_inputs(16)
    rv = np.zeros((16, 4)); rv[:, [0, 2]] = d["x0"][:, [0, 2]]
    tgt = dart_b200.workloads.rmpc_config3(16)["target"]
    refN = problems.build_ref_traj(None, problems.reference_governor(rv, tgt), tgt, N, 4, 0.2)
    out = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(N=N), device=0).solve(d["x0"], refN, aux=d["aux"])
    ref = ipm.solve(problems.rmpc_problem(d["x0"], d["u_prev"], d["theta"], refN, N=N))
    helpers.assert_parity(out, ref, f"rmpc N={N}")


def test_mpc_worker_queue_protocol(built):
    """PMPC/main_parallel.py:10-43: items (state, target) / "STOP"; replies (u_cmd, loss, solve_time)."""
    import queue
    import threading
    sq, cq = queue.Queue(), queue.Queue()
    model, data = dart_b200.GravityModel(-9.81, 0.002), dart_b200.StateHolder()
    params = {"Ts": 0.002, "nx": 6, "nu": 2, "N": 15, "Qp": 600, "Qv": 5, "R": 0.1, "u_bounds": (-0.6, 0.6), "mu": 0.1}
    th = threading.Thread(target=dart_b200.mpc_worker, args=((model, data), "cube", params, sq, cq))
    th.start()
    c, aux, p = helpers.pmpc_case(1)
    cube = [i for i in range(p.B) if c["Qp"][i] == 600 and c["mu"][i] == 0.1][:3]
    for i in cube:
        sq.put((c["state"][i], c["target"][i]))
    sq.put("STOP")
    th.join(timeout=60)
    assert not th.is_alive()
    ref = ipm.solve(p)
    for i in cube:
        u_cmd, loss, solve_time = cq.get(timeout=5)
        assert u_cmd.shape == (2,) and loss.shape == (1,) and solve_time > 0
        assert np.abs(u_cmd - ref["U"][i, 0]).max() <= helpers.TOL_U0
        assert abs(loss[0] - ref["J"][i]) <= helpers.TOL_J * abs(ref["J"][i])


def test_mpc_worker_in_a_spawned_process(built):
    """The reference's launcher contract (PMPC/main_parallel.py:46,138-145, mpc_3d.py:161): the worker runs in a process
    started with the ``spawn`` method and talks through ``mp.Queue`` -- so the arguments and the reply tuple must pickle
    and the CUDA context is created in the child."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    sq, cq = ctx.Queue(), ctx.Queue()
    model, data = dart_b200.GravityModel(-9.81, 0.002), dart_b200.StateHolder()
    params = {"Ts": 0.002, "nx": 6, "nu": 2, "N": 15, "Qp": 600, "Qv": 5, "R": 0.1, "u_bounds": (-0.6, 0.6), "mu": 0.1}
    proc = ctx.Process(target=dart_b200.mpc_worker, args=((model, data), "cube", params, sq, cq), daemon=True)
    proc.start()
    c, aux, p = helpers.pmpc_case(1)
    cube = [i for i in range(p.B) if c["Qp"][i] == 600 and c["mu"][i] == 0.1][:3]
    for i in cube:
        sq.put((c["state"][i], c["target"][i]))
    ref = ipm.solve(p)
    for i in cube:
        u_cmd, loss, solve_time = cq.get(timeout=180)          # the first reply pays the child's imports and CUDA start-up
        assert isinstance(u_cmd, np.ndarray) and u_cmd.shape == (2,) and loss.shape == (1,) and solve_time > 0
        assert np.abs(u_cmd - ref["U"][i, 0]).max() <= helpers.TOL_U0
        assert abs(loss[0] - ref["J"][i]) <= helpers.TOL_J * abs(ref["J"][i])
    sq.put("STOP")
    proc.join(timeout=60)
    assert proc.exitcode == 0


def test_pmpc_class_is_a_dropin(built):
    model, data = dart_b200.GravityModel(-9.81, 0.002), dart_b200.StateHolder()
    ctl = dart_b200.PMPC(model, data, Ts=0.002, nx=6, nu=2, N=15, Qp=400, Qv=2, R=0.2, u_bounds=(-0.6, 0.6), mu=0.1)
    ctl.target_body = "cube"
    data.body("cube").xpos[:] = [0.0, 0.0, 0.43]
    assert np.array_equal(ctl.get_state(), [0, 0, 0, 0, 0.43, 0])
    u0, loss = ctl.solve(np.array([0.1, 0, 0.05, 0, 0.4, 0]))
    assert u0.shape == (2,) and loss.shape == (1,) and ctl.w0.shape == (126,) and ctl.g == -9.81 and ctl.status == 0
    assert np.array_equal(ctl.w0[96:98], u0) and np.all(u0 < 0)        # g < 0: +x / +y motion needs negative tilt


def test_two_devices_in_one_process(built):
    """Handles on different GPUs of one process (kernel attributes are cached per device); wrong current device is refused."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    c, aux, p = helpers.pmpc_case(2)
    outs = []
    for dev in (0, 1):
        torch.cuda.set_device(dev)
        eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=dev)
        outs.append(eng.solve(c["state"], c["target"], aux=aux, want_w=False))
        d = torch.device("cuda", dev)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)
        o = eng.solve_device(t(c["state"]), t(c["target"]), aux=t(aux))
        torch.cuda.synchronize(d)
        assert np.array_equal(o["u0"].cpu().numpy(), outs[-1]["u0"])
        if dev == 1:
            # the C ABI refuses a launch from another current device; the Python wrapper runs under the handle's device
            import ctypes as C
            torch.cuda.set_device(0)
            x, r, a_ = t(c["state"]), t(c["target"]), t(aux)
            u0 = torch.empty((x.shape[0], 2), dtype=torch.float64, device=d); J = torch.empty((x.shape[0],), dtype=torch.float64, device=d)
            p_ = lambda z: C.c_void_p(z.data_ptr())
            rc = dart_b200._lib.lib().dart_solve(eng._h, x.shape[0], p_(x), p_(r), p_(a_), None, None, p_(u0), p_(J), None, None, None)
            assert rc == -1
            o2 = eng.solve_device(x, r, aux=a_)
            torch.cuda.synchronize(d)
            assert torch.cuda.current_device() == 0 and np.array_equal(o2["u0"].cpu().numpy(), outs[-1]["u0"])
    torch.cuda.set_device(0)
    assert np.array_equal(outs[0]["u0"], outs[1]["u0"]) and np.array_equal(outs[0]["J"], outs[1]["J"])


def test_peer_rows_gather_two_gpus(built):
    """Gather without a collective (tests/dist_peer_rows_check.py): the rows the solve kernels store into every rank's buffer
    over NVLink equal the NCCL all_gather bit for bit after the flag hand-shake, for all three methods, over several steps."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29641", os.path.join(helpers.ROOT, "tests", "dist_peer_rows_check.py")],
                       capture_output=True, text=True, timeout=300)
    assert "PEER_ROWS_OK" in r.stdout or "PEER_ROWS_UNAVAILABLE" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
