"""PMPC headline batch under both barrier strategies: time, iterations, agreement of the solutions (run on a B200)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_b200

dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
res = {}
for spo in (64, 7282):
    c = dart_b200.workloads.pmpc_config2(spo, seed=1)
    aux = np.ascontiguousarray(np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], axis=1))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    x0d, tgd, axd = t(c["state"]), t(c["target"]), t(aux)
    B = x0d.shape[0]
    for strat in ("monotone", "mehrotra") + (("mehrotra16",) if spo > 64 else ()):
        eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(lanes=16 if strat == "mehrotra16" else 0), device=0)
        eng.set_barrier_strategy(strat[:8])
        u0 = torch.empty((B, 2), dtype=torch.float64, device=dev); J = torch.empty((B,), dtype=torch.float64, device=dev)
        st = torch.empty((B,), dtype=torch.int32, device=dev); it = torch.empty((B,), dtype=torch.int32, device=dev)
        for _ in range(3):
            eng.solve_device(x0d, tgd, aux=axd, u0_out=u0, J_out=J, status=st, iters=it)
        torch.cuda.synchronize()
        ms = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.solve_device(x0d, tgd, aux=axd, u0_out=u0, J_out=J, status=st, iters=it); b.record()
            torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
        itn = it.cpu().numpy()
        res[(spo, strat)] = (u0.cpu().numpy(), J.cpu().numpy())
        print(f"B={B:7d} {strat:9s} ms {np.median(ms):.4f}  iters mean {itn.mean():.2f} max {itn.max()}  converged {(st == 0).sum().item()}  "
              f"launch {eng.last_launch_config()}", flush=True)
    (ua, Ja), (ub, Jb) = res[(spo, "monotone")], res[(spo, "mehrotra")]
    print(f"   mehrotra vs monotone: max |du0| {np.abs(ua - ub).max():.2e}  max rel dJ {np.abs(Jb / Ja - 1).max():.2e}", flush=True)
