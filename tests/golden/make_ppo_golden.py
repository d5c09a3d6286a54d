"""Generate tests/golden/ppo_reference_checkpoint.npz (run HERE, where /root/reference exists: python tests/golden/make_ppo_golden.py).

The weights are an artefact of the reference itself: the policy it trained and saved with torch.save at rlmpc2.py:917-922
(LMPC/src/checkpoints/general/best_agent.pth["model"]).  Outputs on seeded inputs are computed with torch (the reference's
library) through the restated Policy: rollout-time mean / value / log-probability, the PPO loss terms and the gradient of one
minibatch (per-tensor L2 norms and 64 sampled entries per tensor).  The fixture travels to the GPU box; the checkpoint does not.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dart_b200 import ppo                 # noqa: E402
from oracle import ppo as oppo            # noqa: E402

CKPT = "/root/reference/LMPC/src/checkpoints/general/best_agent.pth"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ppo_reference_checkpoint.npz")
M = 96


def inputs():
    g = torch.Generator().manual_seed(2024)
    obs = torch.randn(M, 520, generator=g)
    eps = torch.randn(M, 34, generator=g)
    dlogp = 0.3 * torch.randn(M, generator=g)
    adv = torch.randn(M, generator=g)
    dret = torch.randn(M, generator=g)
    return obs, eps, dlogp, adv, dret


def compute(flat):
    pol = oppo.Policy()
    pol.load_state_dict({k: torch.from_numpy(v) for k, v in ppo.unpack_params(flat).items()})
    obs, eps, dlogp, adv, dret = inputs()
    action, logp, value, mean = oppo.act(pol, obs, eps)
    old_logp, ret = logp + dlogp, value + dret
    opt = oppo.make_optimizer(pol)
    pl, vl, ent, gn = oppo.minibatch_step(pol, opt, obs, action, old_logp, adv, ret, apply=False)
    rng = np.random.default_rng(7)
    out = dict(mean=mean.numpy(), value=value.numpy(), logp=logp.numpy(), action=action.numpy(), old_logp=old_logp.numpy(),
               adv=adv.numpy(), ret=ret.numpy(), stats=np.array([pl, vl, ent, gn]))
    for k, p in pol.named_parameters():
        gflat = p.grad.numpy().reshape(-1)
        idx = rng.choice(gflat.size, size=min(64, gflat.size), replace=False)
        out["gnorm/" + k] = np.array(np.linalg.norm(gflat.astype(np.float64)))
        out["gidx/" + k] = idx
        out["gval/" + k] = gflat[idx]
    return out


def main():
    ck = torch.load(CKPT, map_location="cpu", weights_only=False)
    flat = ppo.pack_params(ck["model"])
    out = compute(flat)
    np.savez_compressed(OUT, params=flat, episode=np.array(ck.get("episode", -1)), **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; reference episode", ck.get("episode"), "return", ck.get("return"))


if __name__ == "__main__":
    main()
