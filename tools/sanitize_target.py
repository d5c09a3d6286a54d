"""Tiny run of every kernel family for compute-sanitizer (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, dart_b200
W = dart_b200.workloads
c, aux = W.pmpc_inputs(1)
for lanes in (4, 16):
    out = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(lanes=lanes), 0).solve(c["state"][:6], c["target"][:6], aux=aux[:6])
    assert (out["status"] == 0).all()
d = W.rmpc_inputs(3)
for lanes in (8, 32):
    out = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(lanes=lanes), 0).solve(d["x0"], d["ref"], aux=d["aux"])
    assert (out["status"] == 0).all()
d = W.lmpc_inputs(3)
out = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(), 0).solve(d["x0"], d["ref"], aux=d["aux"])
assert (out["status"] == 0).all()
pol = dart_b200.PolicyMLP(seed=3, device=0)
o = pol.forward(torch.randn((300, 520), dtype=torch.float32, device="cuda"))
torch.cuda.synchronize()
print("sanitize target ok", float(o.abs().mean()))
