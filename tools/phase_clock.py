"""Print the per-phase cycle counts of block 0 (library built with -DDART_PHASE_CLOCK; DART_B200_LIB points at it)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import dart_b200
from dart_b200 import workloads as W

def run(method, x0, ref, aux):
    eng = dart_b200.NMPCEngine(getattr(dart_b200, method + "_cfg")(), device=0)
    print("==", method, len(x0), flush=True)
    out = eng.solve(x0, ref, aux=aux)
    torch.cuda.synchronize()
    it = np.asarray(out["iters"])
    print("iters max", int(it.max()), "mean", float(it.mean()), flush=True)

c, aux = W.pmpc_inputs(64)
run("pmpc", c["state"][:1], c["target"][:1], aux[:1])
run("pmpc", c["state"], c["target"], aux)
d = W.rmpc_inputs(4096); run("rmpc", d["x0"], d["ref"], d["aux"])
d = W.lmpc_inputs(16384); run("lmpc", d["x0"], d["ref"], d["aux"])
