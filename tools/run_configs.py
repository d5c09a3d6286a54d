"""BASELINE configs 1, 2 (closed loop), 3, 4 and 5 on the GPU(s): measurement tool, JSON summary on stdout / --out.

  python tools/run_configs.py [--quick] [--out file.json]
  python -m torch.distributed.run --nproc-per-node N ... tools/run_configs.py --only config5     (scale sweep)
bench.py stays the headline contract (config 2, open loop).  Surrogate plants as in SURVEY.md 8(d).
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dart_b200
W = dart_b200.workloads


def ev():
    return torch.cuda.Event(enable_timing=True)


def config1(steps, persistent=False):
    c1 = dart_b200.workloads.pmpc_config1()
    aux = np.stack([c1["Qp"], c1["Qv"], c1["R"], c1["mu"]], axis=1)
    ep = dart_b200.PMPCEpisodes(c1["state"], c1["target"], aux, device=LOCAL)
    if persistent:
        dart_b200.PMPCEpisodes(c1["state"], c1["target"], aux, device=LOCAL).run(10, persistent=True)       # first-launch costs
        t0 = time.perf_counter(); m = ep.run(steps, persistent=True); dt = time.perf_counter() - t0
        m["trace"] = np.zeros((0,))
    else:
        t0 = time.perf_counter(); m = ep.run(steps, trace_every=max(1, steps // 100)); dt = time.perf_counter() - t0
    return dict(config="1: PMPC single episode (cube, mu=0.10, model-as-plant)" + (", all steps in one launch" if persistent else ""), steps=steps, wall_s=dt, solves_per_s=steps / dt,
                ms_per_solve=dt / steps * 1e3, convergence_time_s=float(m["convergence_time"][0]),
                steady_state_error_m=float(m["steady_state_error"][0]), control_effort=float(m["control_effort"][0]),
                mean_iters=m["mean_iters"], not_converged_solves=m["not_converged_solves"], trace=m["trace"].tolist())


def config2_closed_loop(steps, label="2 (closed loop): 18 objects x 64 states, surrogate plant with unmodelled Coulomb term", persistent=False, **kw):
    c, aux = W.pmpc_inputs(64)
    rng = np.random.default_rng(21)
    ep = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, coulomb=rng.uniform(0, 0.02, aux.shape[0]), device=LOCAL, **kw)
    a, b = ev(), ev()
    a.record(); m = ep.run(steps, graph=GRAPH, persistent=persistent); b.record(); torch.cuda.synchronize()
    sec = a.elapsed_time(b) * 1e-3
    per_obj = []
    objs = dart_b200.workloads.pmpc_objects()
    for i, o in enumerate(objs):
        sl = slice(i * 64, (i + 1) * 64)
        per_obj.append(dict(object=f"{o['shape']} m={o['mass']} mu={o['mu']}", settled_frac=float(m["converged"][sl].mean()),
                            convergence_time_s=float(np.median(m["convergence_time"][sl])),
                            steady_state_error_mm=float(np.median(m["steady_state_error"][sl]) * 1e3),
                            control_effort=float(np.median(m["control_effort"][sl]))))
    return dict(config=label, steps=steps, cuda_graph=GRAPH,
                sim_time_s=m["sim_time"], solves=m["solves"], seconds=sec, solves_per_s=m["solves"] / sec,
                mean_iters=m["mean_iters"], not_converged_solves=m["not_converged_solves"], per_object=per_obj)


def config3(B, T, label="3: RMPC + per-instance RLS closed loop", **kw):
    c = dart_b200.workloads.rmpc_config3(B, seed=2)
    dev = torch.device("cuda", LOCAL)
    x = torch.from_numpy(c["x0"]).to(dev)
    ctl = dart_b200.RMPCBatch(B, c["target"], c["x0"], device=LOCAL, **kw)
    rv0 = np.zeros((B, 4)); rv0[:, [0, 2]] = c["x0"][:, [0, 2]]
    ctl.set_virtual_reference(rv0)
    mu = torch.from_numpy(c["mu_plant"]).to(dev); cp = torch.from_numpy(c["c_plant"]).to(dev)
    stat = torch.zeros(4, dtype=torch.int64, device=dev); it_sum = torch.zeros((), dtype=torch.int64, device=dev)
    a, b = ev(), ev()
    a.record()
    for t in range(T):
        u = ctl.step(x)
        stat += torch.bincount(ctl.status.long(), minlength=4)
        it_sum += ctl.iters.sum()
        # surrogate plant: v' = gz sin u - mu g tanh(v/.01) - c v  (semi-implicit Euler, 4 sub-steps; torch = plumbing)
        for _ in range(4):
            h = 0.002 / 4
            ax = -9.81 * torch.sin(u[:, 0]) - mu * 9.81 * torch.tanh(x[:, 1] / 0.01) - cp * x[:, 1]
            ay = -9.81 * torch.sin(u[:, 1]) - mu * 9.81 * torch.tanh(x[:, 3] / 0.01) - cp * x[:, 3]
            x = torch.stack([x[:, 0] + h * x[:, 1], x[:, 1] + h * ax, x[:, 2] + h * x[:, 3], x[:, 3] + h * ay], dim=1).contiguous()
    b.record(); torch.cuda.synchronize()
    sec = a.elapsed_time(b) * 1e-3
    st = stat.cpu().numpy()
    err = (x[:, [0, 2]] - ctl.target[:, [0, 2]]).norm(dim=1)
    return dict(config=label, B=B, steps=T, seconds=sec, solves_per_s=B * T / sec,
                status_counts=dict(converged=int(st[0]), max_iter=int(st[1]), infeasible=int(st[2]), numeric=int(st[3])),
                mean_iters=float(it_sum.item()) / (B * T), median_pos_err_m=float(err.median().item()),
                theta_hat_absmax=float(ctl.theta.abs().max().item()))


def config4(B, T, label="4: LMPC with the policy MLP (random orthogonal init), model-as-plant", **kw):
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    dev = torch.device("cuda", LOCAL)
    ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=3, device=LOCAL, **kw)
    x = torch.from_numpy(c["state"]).to(dev); tg = torch.from_numpy(c["target"]).to(dev)
    stat = torch.zeros(5, dtype=torch.int64, device=dev); it_sum = torch.zeros((), dtype=torch.int64, device=dev)
    a, b = ev(), ev()
    a.record()
    for t in range(T):
        ctl.step(x, tg)
        stat += torch.bincount(ctl.status.long(), minlength=5)
        it_sum += ctl.iters.sum()
        x = ctl.w[:, 8:16].contiguous()          # plant = the controller's own model: predicted x_1 of the optimal plan
    b.record(); torch.cuda.synchronize()
    sec = a.elapsed_time(b) * 1e-3
    st = stat.cpu().numpy()
    return dict(config=label, B=B, steps=T, seconds=sec,
                solves_per_s=B * T / sec, status_counts=dict(converged=int(st[0]), max_iter=int(st[1]), infeasible=int(st[2]), numeric=int(st[3]), acceptable=int(st[4])),
                final_pos_err_m=float((x[:, [0, 2]] - tg[:, [0, 2]]).norm(dim=1).median().item()),
                mean_iters=float(it_sum.item()) / (B * T), policy_launches=ctl.policy.launch_count,
                pvec_range=[float(ctl.pvec.min().item()), float(ctl.pvec.max().item())])


def config5(total):
    """Scale sweep: `total` mixed instances, 1/3 per method, contiguous shards per rank, result rows all-gathered."""
    third = total // 3
    lo, hi = dart_b200.shard_bounds(third, WORLD, RANK)
    n = hi - lo
    dev = torch.device("cuda", LOCAL)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    seed = 100 + RANK
    S = (n + 17) // 18
    cp = dart_b200.workloads.pmpc_config2(S, seed=seed)
    px, pt = t(cp["state"][:n]), t(cp["target"][:n]); pa = t(np.stack([cp["Qp"], cp["Qv"], cp["R"], cp["mu"]], 1)[:n])
    rd = W.rmpc_inputs(n, seed=seed)
    rx, rr, ra = t(rd["x0"]), t(rd["ref"]), t(rd["aux"])
    ld = W.lmpc_inputs(n, seed=seed)
    lx, lr, la = t(ld["x0"]), t(ld["ref"]), t(ld["aux"])
    engs = [dart_b200.NMPCEngine(f(), device=LOCAL) for f in (dart_b200.pmpc_cfg, dart_b200.rmpc_cfg, dart_b200.lmpc_cfg)]
    rows = [torch.empty((n, 4), dtype=torch.float64, device=dev) for _ in range(3)]
    for e, r in zip(engs, rows):
        e.set_result_rows(r)
    outs = [dict(u0_out=torch.empty((n, 2), dtype=torch.float64, device=dev), J_out=torch.empty((n,), dtype=torch.float64, device=dev),
                 status=torch.empty((n,), dtype=torch.int32, device=dev), iters=torch.empty((n,), dtype=torch.int32, device=dev)) for _ in range(3)]
    nmax = (third + WORLD - 1) // WORLD
    pad = torch.zeros((3 * nmax, 4), dtype=torch.float64, device=dev)
    full = torch.empty((WORLD * 3 * nmax, 4), dtype=torch.float64, device=dev) if WORLD > 1 else None

    def step():
        engs[0].solve_device(px, pt, aux=pa, **outs[0])
        engs[1].solve_device(rx, rr, aux=ra, **outs[1])
        engs[2].solve_device(lx, lr, aux=la, **outs[2])
        if WORLD > 1:
            for i in range(3):
                pad[i * nmax: i * nmax + n] = rows[i]
            dist.all_gather_into_tensor(full, pad)

    step(); torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    a, b = ev(), ev()
    a.record(); step(); b.record(); torch.cuda.synchronize()
    sec = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device=dev)
    conv = torch.tensor([float(sum((o["status"] == 0).sum().item() for o in outs))], dtype=torch.float64, device=dev)
    if WORLD > 1:
        dist.all_reduce(sec, op=dist.ReduceOp.MAX); dist.all_reduce(conv, op=dist.ReduceOp.SUM)
    return dict(config="5: scale sweep, mixed PMPC/RMPC/LMPC", total_instances=3 * third, n_gpus=WORLD, seconds=float(sec.item()),
                converged=int(conv.item()), solves_per_s=float(conv.item()) / float(sec.item()),
                per_method_mean_iters=[float(o["iters"].double().mean().item()) for o in outs])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true"); ap.add_argument("--graph", action="store_true"); ap.add_argument("--out", default=None); ap.add_argument("--only", default=None)
    args = ap.parse_args()
    WORLD = int(os.environ.get("WORLD_SIZE", "1")); RANK = int(os.environ.get("RANK", "0")); LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(LOCAL)
    dist = None
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    q = args.quick
    GRAPH = args.graph
    res = []
    todo = args.only.split(",") if args.only else ["config1", "config1_persistent", "config2", "config2_persistent", "config2_warm", "config3", "config3_dual", "config4", "config4_dual", "config4_shift", "config4_refopts", "config5"]
    for name in todo:
        if WORLD > 1 and name != "config5":
            continue
        if name == "config1": r = config1(500 if q else 5000)
        elif name == "config1_persistent": r = config1(500 if q else 5000, persistent=True)
        elif name == "config2_persistent": r = config2_closed_loop(200 if q else 5000, label="2 (f1): closed loop, all 5000 steps in one launch (dart_pmpc_episode)", persistent=True)
        elif name == "config2": r = config2_closed_loop(200 if q else 5000)
        elif name == "config2_warm": r = config2_closed_loop(200 if q else 5000, label="2 (f2): closed loop with primal warm start + warm-started barrier (not the reference's cold start)", warm_start=True)
        elif name == "config3": r = config3(512 if q else 4096, 32 if q else 256)
        elif name == "config3_dual": r = config3(512 if q else 4096, 32 if q else 256, label="3 (f2): dual warm start", dual_warm=True)
        elif name == "config4": r = config4(2048 if q else 16384, 8 if q else 64)
        elif name == "config4_dual": r = config4(2048 if q else 16384, 8 if q else 64, label="4 (f2): dual warm start", dual_warm=True)
        elif name == "config4_shift": r = config4(2048 if q else 16384, 8 if q else 64, label="4 (f2): shifted-plan warm start", warm_start="shift")
        elif name == "config4_refopts":
            from dart_b200.config import LMPC_REFERENCE_SOLVER_OPTIONS as ro
            r = config4(2048 if q else 16384, 8 if q else 64, label="4 (f2): the reference's IPOPT options (tol 1e-4, acceptable 1e-3 x 5, max_iter 50) + plan-shift fallback", plan_fallback=True, **ro)
        else: r = config5(3 * 2 ** 12 if q else 2 ** 20)
        res.append(r)
        if RANK == 0:
            print(json.dumps({k: v for k, v in r.items() if k not in ("trace", "per_object")})[:1200], flush=True)
    if RANK == 0 and args.out:
        json.dump(res, open(args.out, "w"), indent=1)
    if WORLD > 1:
        dist.destroy_process_group()
