import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dart_b200
dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for S in (64, 7282):
    c = dart_b200.workloads.pmpc_config2(S, seed=1)
    B = c["state"].shape[0]
    x, tg, ax = t(c["state"]), t(c["target"]), t(np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], 1))
    for lanes in (8, 16):
        for bt in (0, 32, 64, 128, 256):
            try:
                eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(lanes=lanes, block_threads=bt), device=0)
                o = dict(u0_out=torch.empty((B, 2), dtype=torch.float64, device=dev), J_out=torch.empty((B,), dtype=torch.float64, device=dev),
                         status=torch.empty((B,), dtype=torch.int32, device=dev), iters=torch.empty((B,), dtype=torch.int32, device=dev))
                for _ in range(3):
                    eng.solve_device(x, tg, aux=ax, **o)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 20 if S == 64 else 3
                a.record()
                for _ in range(reps):
                    eng.solve_device(x, tg, aux=ax, **o)
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / reps
                print(json.dumps(dict(B=B, lanes=lanes, bt=bt, ms=round(ms, 4), Msolves=round(B / ms / 1e3, 2), conv=int((o["status"] == 0).sum()), iters=float(o["iters"].double().mean()), cfg=eng.last_launch_config())), flush=True)
                eng.close()
            except Exception as e:
                print(json.dumps(dict(B=B, lanes=lanes, bt=bt, error=str(e)[:80])))
