"""Where the host-API (e2e) time goes for the 1152-instance batch (dev tool)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dart_b200
c, aux = dart_b200.workloads.pmpc_inputs(64)
eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), 0)
dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
x, tg, ax = t(c["state"]), t(c["target"]), t(aux)
def med(f, n=300):
    for _ in range(20): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return np.median(ts) * 1e6
host = med(lambda: eng.solve(c["state"], c["target"], aux=aux, want_w=False))
def devsync():
    eng.solve_device(x, tg, aux=ax); torch.cuda.synchronize()
dv = med(devsync)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100): eng.solve_device(x, tg, aux=ax)
b.record(); torch.cuda.synchronize()
print(f"host API p50 {host:.1f} us | device API + sync p50 {dv:.1f} us | kernel (events, back-to-back) {a.elapsed_time(b) * 10:.1f} us")
pin = torch.empty((1152, 16), dtype=torch.float64).pin_memory(); dd = torch.empty((1152, 16), dtype=torch.float64, device=dev)
def h2d():
    dd.copy_(pin, non_blocking=True); torch.cuda.synchronize()
print(f"pinned H2D 147 KB + sync p50 {med(h2d):.1f} us")
src = np.zeros((1152, 16)); dst = np.zeros((1152, 16))
print(f"host memcpy 147 KB p50 {med(lambda: np.copyto(dst, src)):.1f} us")
