"""Timing of every kernel family on the BASELINE configs 3-5 shapes (dev/measurement tool, CUDA events).

  python tools/bench_methods.py [--sweep]
Prints one JSON object per measurement; bench.py stays the headline contract (config 2).
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dart_b200
W = dart_b200.workloads

DEV = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
# SURVEY 8(d), per IPM iteration (+ the corrector's vector sweeps where predictor-corrector steps run: bench.py has the derivation)
FLOPS = {"pmpc": 66.9e3, "rmpc": 55.3e3, "lmpc": 245.7e3}
if not os.environ.get("DART_BARRIER_MONOTONE"):
    FLOPS = {"pmpc": 66.9e3 + 4.68e3, "rmpc": 55.3e3 + 6.24e3, "lmpc": 245.7e3 + 15.2e3}


def time_solve(method, cfg, x0, ref, aux, reps=5):
    eng = dart_b200.NMPCEngine(cfg, 0)
    B = x0.shape[0]
    u0 = torch.empty((B, 2), dtype=torch.float64, device=DEV); J = torch.empty((B,), dtype=torch.float64, device=DEV)
    st = torch.empty((B,), dtype=torch.int32, device=DEV); it = torch.empty((B,), dtype=torch.int32, device=DEV)
    X, R, A = t(x0), t(ref), (None if aux is None else t(aux))
    for _ in range(2):
        eng.solve_device(X, R, aux=A, u0_out=u0, J_out=J, status=st, iters=it)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.solve_device(X, R, aux=A, u0_out=u0, J_out=J, status=st, iters=it)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    its = int(it.sum().item()); ok = int((st == 0).sum().item())
    return dict(method=method, B=B, ms=round(ms, 4), solves_per_s=ok / ms * 1e3, converged=ok, mean_iters=its / B, max_iters=int(it.max().item()),
                fp64_tflops=its * FLOPS[method] / (ms * 1e-3) / 1e12, launch=eng.last_launch_config())


def tile_rows(a, reps):
    return np.tile(a, (reps,) + (1,) * (a.ndim - 1))


def main():
    sweep = "--sweep" in sys.argv
    peak = dart_b200.measure_fp64_tflops(0)
    print(json.dumps({"fp64_peak_tflops_measured": peak}))
    # config 3: RMPC 4096 instances (mid-episode inputs)
    d = W.rmpc_inputs(4096)
    for lanes in ((8, 16, 32) if sweep else (0,)):
        for bt in ((32, 64, 128) if sweep else (0,)):
            try:
                print(json.dumps(time_solve("rmpc", dart_b200.rmpc_cfg(lanes=lanes, block_threads=bt), d["x0"], d["ref"], d["aux"])), flush=True)
            except Exception as e:
                print(json.dumps({"method": "rmpc", "lanes": lanes, "bt": bt, "error": str(e)[:100]}))
    # config 4: LMPC 16384 instances
    d = W.lmpc_inputs(16384)
    for lanes in ((8, 16, 32) if sweep else (0,)):
        for bt in ((32, 64, 128) if sweep else (0,)):
            try:
                print(json.dumps(time_solve("lmpc", dart_b200.lmpc_cfg(lanes=lanes, block_threads=bt), d["x0"], d["ref"], d["aux"], reps=3)), flush=True)
            except Exception as e:
                print(json.dumps({"method": "lmpc", "lanes": lanes, "bt": bt, "error": str(e)[:100]}))
    # policy MLP: HBM roofline (2080 B in + 136 B out per instance)
    pol = dart_b200.PolicyMLP(seed=3, device=0)
    hbm = 6552.0
    try:
        hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for B in (16384, 262144, 1048576):
        obs = torch.randn((B, 520), dtype=torch.float32, device=DEV); out = torch.empty((B, 34), dtype=torch.float32, device=DEV)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
        for _ in range(3):
            pol.forward(obs, out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); pol.forward(obs, out); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        gbs = B * (2080 + 136) / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": "policy_mlp", "B": B, "ms": round(ms, 4), "GBps": gbs, "frac_of_measured_hbm": gbs / hbm,
                          "tflops_tf32": B * 79104 / (ms * 1e-3) / 1e12}), flush=True)
    # RLS prologue
    B = 4096
    ctl = dart_b200.RMPCBatch(B, np.zeros((B, 4)), np.zeros((B, 4)), device=0)
    x = torch.zeros((B, 4), dtype=torch.float64, device=DEV)
    for _ in range(3):
        ctl.step(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ctl.step(x)
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"loop": "RMPCBatch.step (prologue + warm solve)", "B": B, "ms": a.elapsed_time(b) / 10,
                      "mean_iters": float(ctl.iters.double().mean().item())}))


if __name__ == "__main__":
    main()
