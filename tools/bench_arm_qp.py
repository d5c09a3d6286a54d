"""Throughput of the batched arm QP on one GPU: kernel-only (inputs resident) and through ArmQPBatch
(H2D + device QP build + solve + D2H).  JSON lines on stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dart_b200
from dart_b200 import arm as parm

dev = torch.device("cuda", 0)
P = parm.default_params()
for B, stress in ((2304, 1.0), (32768, 1.0), (262144, 1.0), (262144, 0.3), (262144, 3.0)):
    dyn = dart_b200.workloads.arm_dynamics(B, seed=7, stress=stress)
    H, g, c0, C, lo, hi = parm.build_qp(dyn, P)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    Hd, gd, Cd, lod, hid = t(H), t(g), t(C), t(lo), t(hi)
    out = parm.solve_qp_device(Hd, gd, Cd, lod, hid)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); parm.solve_qp_device(Hd, gd, Cd, lod, hid, out=out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    byts = B * (49 + 7 + 147 + 42 + 7 + 1) * 8 + B * 8
    # per iteration and QP: 21x7 rows x 2 (residual, step) + 49 (Hx) + 2x147 (C'nu, C'nuhat) + 21x7x7 (C'SC) + ~170 (factor/solve)
    flops = 2.0 * float(it.sum()) * (2 * 147 + 49 + 2 * 147 + 1029 + 170)
    print(json.dumps(dict(kernel="arm_qp_kernel", B=B, stress=stress, ms=round(ms, 4), qps_per_s=B / ms * 1e3, converged=int((st == 0).sum()),
                          mean_iters=float(it.mean()), max_iters=int(it.max()), hbm_GBps=byts / ms * 1e-6, fp64_tflops=flops / ms * 1e-9)), flush=True)
ctl = dart_b200.ArmQPBatch(P, device=0)
dyn = dart_b200.workloads.arm_dynamics(2304, seed=7, stress=1.0)
ctl.solve(dyn)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); ctl.solve(dyn); ts.append(time.perf_counter() - t0)
t0 = time.perf_counter(); parm.build_qp(dyn, P); tb = time.perf_counter() - t0
print(json.dumps(dict(path="ArmQPBatch.solve (H2D of the MuJoCo quantities + device QP build + solve + D2H)", B=2304, ms=float(np.median(ts)) * 1e3,
                      host_numpy_build_ms_for_comparison=tb * 1e3, qps_per_s=2304 / float(np.median(ts)))), flush=True)
