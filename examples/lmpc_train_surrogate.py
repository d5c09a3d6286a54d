#!/usr/bin/env python
"""Batched closed-loop PPO training of the LMPC parameter-adaptation policy on the surrogate plant.

What LMPC/src/run.py does with ``--train`` for ONE controller in MuJoCo (RLMPC with params["train"] = True: the RL worker of
rlmpc2.py:536-935 samples an action every control step, records a transition every 8th step and runs the PPO update block every
``rollout_len`` transitions) for B controllers that share one policy, entirely on the device: observation push -> actor/critic
forward + sample -> parameter update -> NLP solve -> plant step -> reward, then ``epochs`` x minibatch updates on the pooled
[rollout_len, B] rollout.  The plant is the controller's own 8-state model evaluated with per-instance TRUE parameters that the
controller does not know; episodes that end (out of bounds / step cap) restart from their initial state.

    python examples/lmpc_train_surrogate.py --instances 4096 --rollouts 8 --checkpoint /tmp/lmpc_agent.pth
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_b200                                              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=1024)
    ap.add_argument("--rollouts", type=int, default=4)
    ap.add_argument("--rollout-len", type=int, default=16, help="transitions per instance and rollout (one per 8 control steps)")
    ap.add_argument("--mini-batch-size", type=int, default=4096)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--checkpoint", default=None, help="write a checkpoint in the reference's format here when done")
    ap.add_argument("--seed", type=int, default=3)
    a = ap.parse_args()
    import torch
    B = a.instances
    c = dart_b200.workloads.lmpc_config4(B, seed=a.seed)
    rng = np.random.default_rng(a.seed + 1)
    true_pvec = np.clip(c["pvec"] + 0.3 * rng.standard_normal((B, 34)), 0.05, 1.8)
    true_aux = torch.from_numpy(np.concatenate([np.zeros((B, 2)), true_pvec], axis=1)).cuda()
    ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=a.seed)                       # run.py:118-151 controller parameters
    ppo = dart_b200.PPOTrainer(capacity=max(B, a.mini_batch_size), seed=a.seed, lr=a.lr, epochs=a.epochs,
                               mini_batch_size=a.mini_batch_size, vf_coef=0.5,
                               reward_cfg=dict(max_delta=0.02, w_pos=40.0, w_d_ctrl=10.0, max_episode_steps=20000))
    gen = torch.Generator(device="cuda").manual_seed(a.seed)
    loop = dart_b200.LMPCTrainer(ctl, ppo, rollout_len=a.rollout_len, record_every=8, generator=gen)
    x0 = torch.from_numpy(c["state"]).cuda(); target = torch.from_numpy(c["target"]).cuda()
    x = x0.clone()
    steps = a.rollouts * a.rollout_len * 8
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(steps):
        u0, reward, done = loop.step(x, target)
        x = dart_b200.lmpc_plant_step(x, u0, true_aux)
        x = torch.where(done[:, None] > 0, x0, x)
        if (k + 1) % (a.rollout_len * 8) == 0:
            torch.cuda.synchronize()
            err = float((x[:, [0, 2]] - target[:, [0, 2]]).norm(dim=1).median())
            print(f"rollout {len(loop.mean_reward):3d}: mean reward {loop.mean_reward[-1]:9.3f}  optimiser steps {loop.updates:5d}  "
                  f"median position error {err * 1e3:7.2f} mm  {B * (k + 1) / (time.perf_counter() - t0) / 1e6:.2f} M control steps/s")
    if a.checkpoint:
        ppo.save(a.checkpoint, episode=len(loop.mean_reward))
        print("checkpoint written:", a.checkpoint, "(its 'model' entry loads into the reference's Policy)")


if __name__ == "__main__":
    main()
