"""Closed-loop PPO training of the LMPC policy on the surrogate plant (SURVEY 8f.4 + 8d config 4 shape): B controllers with
unknown true parameters, shared policy, rollout of T transitions per instance, then the reference's epochs x minibatches.
Reports wall time per phase and the mean reward per rollout.  Lives under tests/ beside perf_ppo.py; JSON on stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dart_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ROLLOUTS = int(sys.argv[2]) if len(sys.argv) > 2 else 6
T, EVERY, MB, EPOCHS = 16, 8, 4096, 4
c = dart_b200.workloads.lmpc_config4(B, seed=3)
rng = np.random.default_rng(1)
true_aux = torch.from_numpy(np.concatenate([np.zeros((B, 2)), np.clip(c["pvec"] + 0.3 * rng.standard_normal((B, 34)), 0.05, 1.8)], axis=1)).cuda()
ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=3)
ppo = dart_b200.PPOTrainer(capacity=max(B, MB), epochs=EPOCHS, mini_batch_size=MB, lr=3e-4,
                           reward_cfg=dict(max_delta=0.02, w_pos=40.0, max_episode_steps=20000))
g = torch.Generator(device="cuda").manual_seed(7)
tr = dart_b200.LMPCTrainer(ctl, ppo, rollout_len=T, record_every=EVERY, generator=g)
x0 = torch.from_numpy(c["state"]).cuda(); tg = torch.from_numpy(c["target"]).cuda()
x = x0.clone()
torch.cuda.synchronize(); t0 = time.perf_counter()
resets = 0
for k in range(ROLLOUTS * T * EVERY):
    u0, rew, done = tr.step(x, tg)
    x = dart_b200.lmpc_plant_step(x, u0, true_aux)
    x = torch.where(done[:, None] > 0, x0, x)            # reset finished episodes to their initial state
torch.cuda.synchronize(); sec = time.perf_counter() - t0
steps = ROLLOUTS * T * EVERY
print(json.dumps(dict(path="LMPCTrainer closed loop (obs push + actor/critic forward + sample + param update + NLP solve + plant + reward; PPO update per rollout)",
                      B=B, control_steps=steps, rollouts=ROLLOUTS, transitions_per_rollout=T * B, optimiser_steps=tr.updates,
                      seconds=sec, control_steps_per_s=B * steps / sec, mean_reward_per_rollout=tr.mean_reward,
                      ppo_launches=ppo.launch_count, final_pos_err_m=float((x[:, [0, 2]] - tg[:, [0, 2]]).norm(dim=1).median()))))
