import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dart_b200
from oracle import ppo as oppo

from tests.test_gpu_ppo import _rollout, _perturbed_policy
for M in (1000, 16384):
    pol = _perturbed_policy()
    obs, eps, action, logp, value, mean, old_logp, adv, ret = _rollout(M, M + 7, pol)
    opt = oppo.make_optimizer(pol)
    oppo.minibatch_step(pol, opt, obs, action, old_logp, adv, ret, apply=False)
    # float64 reference of the same gradient
    pol64 = _perturbed_policy().double()
    opt64 = oppo.make_optimizer(pol64)
    try:
        oppo.minibatch_step(pol64, opt64, obs.double(), action.double(), old_logp.double(), adv.double(), ret.double(), apply=False)
        have64 = True
    except Exception as e:
        have64 = False; print("no f64 ref", e)
    for mode in ("tc", "simt"):
        if mode == "simt": os.environ["DART_PPO_SIMT"] = "1"
        else: os.environ.pop("DART_PPO_SIMT", None)
        tr = dart_b200.PPOTrainer(capacity=M, state_dict=pol.state_dict())
        tr.update_minibatch(obs.cuda(), action.cuda(), old_logp.cuda(), adv.cuda(), ret.cuda(), apply=False)
        grad = tr.gradient(); tr.close()
        for k, p in pol.named_parameters():
            if not k.endswith("0.weight") and not k.endswith("0.bias"): continue
            ref = p.grad.numpy().astype(np.float64)
            r2 = np.linalg.norm(grad[k] - ref) / np.linalg.norm(ref)
            msg = f"M={M} {mode} {k}: rel2 vs torch32 {r2:.2e} max {np.abs(grad[k]-ref).max()/np.abs(ref).max():.2e}"
            if have64:
                r64 = dict(pol64.named_parameters())[k].grad.numpy()
                msg += f" | vs f64 {np.linalg.norm(grad[k] - r64) / np.linalg.norm(r64):.2e} (torch32 vs f64 {np.linalg.norm(ref - r64) / np.linalg.norm(r64):.2e})"
            print(msg)
