"""Fixed cost and per-iteration cost of the PMPC solve kernel at the headline batch size: the config-2 batch, and the same
number of copies of its fastest / slowest instance (run on a B200: python tools/iter_cost.py)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_b200

dev = torch.device("cuda", 0)
eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=0)
c = dart_b200.workloads.pmpc_config2(64, seed=1)
aux = np.ascontiguousarray(np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], axis=1))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def run(x0, tg, ax, label):
    B = x0.shape[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    x0d, tgd, axd = t(x0), t(tg), t(ax)
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev); J = torch.empty((B,), dtype=torch.float64, device=dev)
    st = torch.empty((B,), dtype=torch.int32, device=dev); it = torch.empty((B,), dtype=torch.int32, device=dev)
    for _ in range(3):
        eng.solve_device(x0d, tgd, aux=axd, u0_out=u0, J_out=J, status=st, iters=it)
    ms = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.solve_device(x0d, tgd, aux=axd, u0_out=u0, J_out=J, status=st, iters=it); b.record()
        torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    itn = it.cpu().numpy()
    print(f"{label:40s} B={B:6d} ms {np.median(ms):.4f}  iters mean {itn.mean():.2f} max {itn.max()}  conv {(st == 0).sum().item()}", flush=True)
    return itn


itn = run(c["state"], c["target"], aux, "config 2")
order = np.argsort(itn)
rep = lambda i, n: (np.repeat(c["state"][i:i + 1], n, 0), np.repeat(c["target"][i:i + 1], n, 0), np.repeat(aux[i:i + 1], n, 0))
for i in (order[0], order[len(order) // 2], order[-1], order[-2], order[-3]):
    run(*rep(i, 1152), f"1152 copies of instance {i} ({itn[i]} it)")
    run(*rep(i, 1), f"1 copy of instance {i} ({itn[i]} it)")
# the batch without its slowest instances
for cut in (14, 13, 12, 11):
    keep = itn <= cut
    run(c["state"][keep], c["target"][keep], aux[keep], f"config 2, instances with <= {cut} it")
