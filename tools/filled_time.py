"""Filled-GPU PMPC batch (131 076 instances): kernel time of the library DART_B200_LIB points at (dev tool for A/B builds)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_b200
dev = torch.device("cuda", 0)
c = dart_b200.workloads.pmpc_config2(7282, seed=1)
aux = np.ascontiguousarray(np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], axis=1))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
x0d, tgd, axd = t(c["state"]), t(c["target"]), t(aux)
B = x0d.shape[0]
eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=0)
u0 = torch.empty((B, 2), dtype=torch.float64, device=dev); J = torch.empty((B,), dtype=torch.float64, device=dev)
st = torch.empty((B,), dtype=torch.int32, device=dev); it = torch.empty((B,), dtype=torch.int32, device=dev)
for _ in range(3):
    eng.solve_device(x0d, tgd, aux=axd, u0_out=u0, J_out=J, status=st, iters=it)
torch.cuda.synchronize()
ms = []
for _ in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.solve_device(x0d, tgd, aux=axd, u0_out=u0, J_out=J, status=st, iters=it); b.record()
    torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
print(os.environ.get("DART_B200_LIB", "production"), f"ms {np.median(ms):.4f} converged {(st == 0).sum().item()} iters {it.float().mean().item():.2f}", eng.last_launch_config(), flush=True)
