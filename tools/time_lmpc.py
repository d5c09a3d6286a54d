"""LMPC solve time at the BASELINE config-4 batch (dev tool for A/B builds through DART_B200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, dart_b200
W = dart_b200.workloads
d = W.lmpc_inputs(16384)
eng = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(), 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
x, r, a = t(d["x0"]), t(d["ref"]), t(d["aux"])
out = eng.solve_device(x, r, aux=a)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.solve_device(x, r, aux=a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.environ.get("DART_B200_LIB", "default"), "lmpc 16384: %.3f ms" % np.median(ts), "converged", int((out["status"] == 0).sum()))
