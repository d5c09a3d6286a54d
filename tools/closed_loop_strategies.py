"""Closed loops of configs 3 / 4 under both barrier strategies and a range of warm-start multiplier scales (dev tool, GPU):
mean iterations per solve, time per step, statuses.  python tools/closed_loop_strategies.py [rmpc|lmpc] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dart_b200
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 48


def run_rmpc(strategy, warm_mu, B=4096):
    c = dart_b200.workloads.rmpc_config3(B, seed=2)
    x = torch.from_numpy(c["x0"]).to(dev)
    ctl = dart_b200.RMPCBatch(B, c["target"], c["x0"], device=0, warm_mu=warm_mu)
    ctl.engine.set_barrier_strategy(strategy)
    rv0 = np.zeros((B, 4)); rv0[:, [0, 2]] = c["x0"][:, [0, 2]]
    ctl.set_virtual_reference(rv0)
    mu = torch.from_numpy(c["mu_plant"]).to(dev); cp = torch.from_numpy(c["c_plant"]).to(dev)
    its = 0; bad = 0; ms = []
    for t in range(T):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); u = ctl.step(x); b.record(); torch.cuda.synchronize()
        if t > 0:
            its += int(ctl.iters.sum().item()); ms.append(a.elapsed_time(b)); bad += int((ctl.status != 0).sum().item())
        x = dart_b200.rmpc_plant_step_device(x, u, mu, cp)
    return its / (B * (T - 1)), float(np.mean(ms)), bad, float(x.abs().sum().item())


def run_lmpc(strategy, warm_mu, B=16384):
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=3, device=0, warm_mu=warm_mu)
    ctl.engine.set_barrier_strategy(strategy)
    x = torch.from_numpy(c["state"]).to(dev); tg = torch.from_numpy(c["target"]).to(dev)
    its = 0; bad = 0; ms = []
    for t in range(T):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ctl.step(x, tg); b.record(); torch.cuda.synchronize()
        if t > 0:
            its += int(ctl.iters.sum().item()); ms.append(a.elapsed_time(b)); bad += int(((ctl.status != 0) & (ctl.status != 4)).sum().item())
        x = ctl.w[:, 8:16].contiguous()
    return its / (B * (T - 1)), float(np.mean(ms)), bad, float(x.abs().sum().item())


for name, fn in (("rmpc", run_rmpc), ("lmpc", run_lmpc)):
    if which not in (name, "both"):
        continue
    for strategy in ("monotone", "mehrotra"):
        for wm in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6):
            it, ms, bad, chk = fn(strategy, wm)
            print(f"{name} {strategy:9s} warm_mu {wm:.0e}: {it:.2f} iterations per warm solve, {ms:.3f} ms per step, {bad} not converged, state checksum {chk:.9g}", flush=True)
