"""Policy MLP forward: accuracy against torch fp64 / fp32 and streaming rate, for both arithmetic modes (measurement tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dart_b200
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
w = dart_b200.init_policy_weights(seed=3)
w = [(W, (0.05 * rng.standard_normal(b.shape)).astype(np.float32)) for W, b in w]


def ref(obs, dt):
    h = torch.from_numpy(obs).to(dt)
    for i, (W, b) in enumerate(w):
        h = h @ torch.from_numpy(W).to(dt).T + torch.from_numpy(b).to(dt)
        if i < 2:
            h = torch.tanh(h)
    return h.numpy()


for prec in ("fp32", "tf32"):
    pol = dart_b200.PolicyMLP(w, device=0, precision=prec)
    for B in (1, 127, 1000, 16384):
        obs = rng.standard_normal((B, 520)).astype(np.float32)
        out = pol.forward(torch.from_numpy(obs).to(dev)).cpu().numpy()
        r64, r32 = ref(obs, torch.float64), ref(obs, torch.float32)
        print(json.dumps(dict(precision=prec, B=B, err_vs_fp64=float(np.abs(out - r64).max()), torch_fp32_vs_fp64=float(np.abs(r32 - r64).max()))), flush=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for B in (16384, 65536, 262144, 1048576):
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
        ms = float(np.median(ts))
        print(json.dumps(dict(precision=prec, B=B, ms=round(ms, 4), GBps=round(B * 2216 / ms / 1e6, 1), frac_of_6552=round(B * 2216 / ms / 1e6 / 6552, 3))), flush=True)
