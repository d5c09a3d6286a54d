import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import dart_b200
from dart_b200 import workloads as W
c, aux = W.pmpc_inputs(64)
eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=0)
out = eng.solve(c["state"], c["target"], aux=aux)
torch.cuda.synchronize()
print("iters", out["iters"].max(), out["iters"].mean(), out["iters"][:4], eng.last_launch_config())
