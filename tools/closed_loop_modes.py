"""Closed-loop PMPC suite (config 2) in four modes: cold/warm start x eager/CUDA-graph replay (dev tool)."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, dart_b200
W = dart_b200.workloads
c, aux = W.pmpc_inputs(64)
rng = np.random.default_rng(21)
cou = rng.uniform(0, 0.02, aux.shape[0])
ep = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, coulomb=cou, device=0)
ep.run(50, persistent=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); m = ep.run(5000, persistent=True); b.record(); torch.cuda.synchronize()
sec = a.elapsed_time(b) * 1e-3
print(f"persistent (one launch for 5000 steps): {sec:.3f} s, {5000*1152/sec/1e6:.1f} M solves/s, mean iters {m['mean_iters']:.2f}, not converged {m['not_converged_solves']}, final err {np.median(m['steady_state_error']):.2e}")
c1 = dart_b200.workloads.pmpc_config1() if hasattr(dart_b200.workloads, 'pmpc_config1') else None
for warm in (False, True):
    for graph in (False, True):
        ep = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, coulomb=cou, device=0, warm_start=warm)
        ep.run(50, graph=graph)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m = ep.run(5000, graph=graph); b.record(); torch.cuda.synchronize()
        sec = a.elapsed_time(b) * 1e-3
        print(f"warm={warm} graph={graph}: {sec:.3f} s, {5000*1152/sec/1e6:.1f} M solves/s, mean iters {m['mean_iters']:.2f}, not converged {m['not_converged_solves']}, final err {np.median(m['steady_state_error']):.2e}")
