"""Timing sweep of the PMPC solve kernel over lanes-per-problem and block size (dev tool, CUDA events)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dart_b200

def timeit(cfg, c, reps=20):
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    eng = dart_b200.NMPCEngine(cfg, 0)
    aux = np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], 1)
    x, tg, ax = t(c["state"]), t(c["target"]), t(aux)
    B = x.shape[0]
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev); J = torch.empty((B,), dtype=torch.float64, device=dev)
    st = torch.empty((B,), dtype=torch.int32, device=dev); it = torch.empty((B,), dtype=torch.int32, device=dev)
    for _ in range(3):
        eng.solve_device(x, tg, aux=ax, u0_out=u0, J_out=J, status=st, iters=it)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        eng.solve_device(x, tg, aux=ax, u0_out=u0, J_out=J, status=st, iters=it)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return ms, int((st == 0).sum()), eng.last_launch_config()

if __name__ == "__main__":
    for S in (64, 7282):
        c = dart_b200.workloads.pmpc_config2(S)
        for lanes in (2, 4, 8, 16):
            for bt in (32, 64, 128, 256):
                try:
                    ms, ok, lc = timeit(dart_b200.pmpc_cfg(lanes=lanes, block_threads=bt), c, 20 if S == 64 else 3)
                    print(f"B={18*S} lanes={lanes} bt={bt} ms={ms:.4f} solves/s={18*S/ms*1e3:.3e} ok={ok} grid={lc['grid']} smem={lc['smem_bytes']}", flush=True)
                except Exception as e:
                    print(f"B={18*S} lanes={lanes} bt={bt} failed: {e}", flush=True)
