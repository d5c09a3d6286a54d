"""Small driver for ncu: a few policy-MLP launches at B = 262144 (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dart_b200
pol = dart_b200.PolicyMLP(seed=3, device=0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
obs = torch.randn((B, 520), dtype=torch.float32, device="cuda"); out = torch.empty((B, 34), dtype=torch.float32, device="cuda")
for _ in range(6):
    pol.forward(obs, out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
