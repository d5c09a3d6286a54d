"""Iteration counts of barrier-update variants on the headline batch (CPU, oracle IPM; python tools/ipm_variants.py [states]).

Why: the headline launch lasts as long as its slowest instance (15 iterations against a mean of 9).  The slow instances hold
a weakly active input bound whose pair (slack, multiplier) must shrink together; a Newton step on z s = mu then only quarters
the product per iteration (both factors halve), so every barrier reduction costs 3-5 iterations.  A second-order corrector
(one more solve with the same factorisation) removes that; Mehrotra's predictor additionally picks mu per iteration."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_b200
from oracle import ipm, problems

spo = int(sys.argv[1]) if len(sys.argv) > 1 else 64
c = dart_b200.workloads.pmpc_config2(spo, seed=1)
p = problems.pmpc_problem(c["state"], c["target"], Qp=c["Qp"], Qv=c["Qv"], R=c["R"], mu=c["mu"])
ref = ipm.solve(p, opts=ipm.Options(tol=1e-10))
for name, kw in (("monotone (the solver's method)", {}), ("monotone + corrector", dict(corrector=1)), ("Mehrotra predictor-corrector", dict(mehrotra=1))):
    t = time.time()
    o = ipm.solve(p, opts=ipm.Options(**kw))
    it = o["iters"]
    print(f"{name:32s} iterations mean {it.mean():.2f} max {it.max()}  converged {(o['status'] == 0).sum()}/{len(it)}  "
          f"|du0| vs the 1e-10 solution {np.abs(o['U'][:, 0] - ref['U'][:, 0]).max():.1e}  rel dJ {np.abs(o['J'] / ref['J'] - 1).max():.1e}  ({time.time() - t:.0f} s)")
