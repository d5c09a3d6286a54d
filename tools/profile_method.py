"""Small driver for ncu: a few solves of one method at its BASELINE batch size (dev tool). usage: profile_method.py pmpc|rmpc|lmpc"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, dart_b200
W = dart_b200.workloads
m = sys.argv[1]
if m == "pmpc":
    c, aux = W.pmpc_inputs(64); d = dict(x0=c["state"], ref=c["target"], aux=aux); cfg = dart_b200.pmpc_cfg()
elif m == "rmpc":
    d = W.rmpc_inputs(4096); cfg = dart_b200.rmpc_cfg()
else:
    d = W.lmpc_inputs(16384); cfg = dart_b200.lmpc_cfg()
eng = dart_b200.NMPCEngine(cfg, 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
x, r, a = t(d["x0"]), t(d["ref"]), t(d["aux"])
for _ in range(4):
    out = eng.solve_device(x, r, aux=a)
torch.cuda.synchronize()
print("ok", int((out["status"] == 0).sum()), eng.last_launch_config())
