import sys, os, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import dart_b200
dev = torch.device("cuda", 0)
w = dart_b200.init_policy_weights(seed=3)
pol = dart_b200.PolicyMLP(w, device=0, precision="fp32")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for B in [int(x) for x in (sys.argv[1:] or ["16384", "1048576"])]:
    obs = torch.randn((B, 520), dtype=torch.float32, device=dev); out = torch.empty((B, 34), dtype=torch.float32, device=dev)
    for _ in range(3):
        pol.forward(obs, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pol.forward(obs, out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(os.environ.get("DART_B200_LIB", "prod"), B, round(float(np.median(ts)), 4), flush=True)
