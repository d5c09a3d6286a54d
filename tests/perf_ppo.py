"""PPO minibatch step on one GPU (lives under tests/ because it times the oracle as the CPU baseline): device time per optimiser step against torch on the host cores (the reference's
own library on its CPU path) for the reference's minibatch (64) and for pooled minibatches.  JSON lines on stdout.

Algorithmic flops per sample: forward 2*(520*128 + 2*64*64 + 64*35) = 153 984, weight gradients the same, data gradients
2*(2*64*64 + 64*35) = 20 864  ->  328 832 flop per sample and step.  Algorithmic HBM bytes per sample: 2 080 (observation,
read by the forward and by the layer-1 weight gradient: 2 x when it does not stay in L2) + 136 (action) + 12."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dart_b200
from oracle import ppo as oppo

FLOP_PER_SAMPLE = 2 * (520 * 128 + 2 * 64 * 64 + 64 * 35) * 2 + 2 * (2 * 64 * 64 + 64 * 35)
dev = torch.device("cuda", 0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
SIZES = [int(a) for a in sys.argv[1:] if a.isdigit()] or [64, 1024, 16384, 65536]
NOCPU = "nocpu" in sys.argv
for M in SIZES:
    g = torch.Generator().manual_seed(M)
    pol = oppo.make_policy(3)
    obs, eps = torch.randn(M, 520, generator=g), torch.randn(M, 34, generator=g)
    act, logp, val, _ = oppo.act(pol, obs, eps)
    old, adv, ret = logp + 0.1 * torch.randn(M, generator=g), torch.randn(M, generator=g), val + torch.randn(M, generator=g)
    tr = dart_b200.PPOTrainer(capacity=M, state_dict=pol.state_dict())
    d = [t.cuda() for t in (obs, act, old, adv, ret)]
    for _ in range(3):
        tr.update_minibatch(*d)
    l0 = tr.launch_count
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); tr.update_minibatch(*d); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    launches = (tr.launch_count - l0) // 7
    # back-to-back steps (no flush): what an epoch loop sees
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        tr.update_minibatch(*d)
    b.record(); torch.cuda.synchronize()
    ms_b2b = a.elapsed_time(b) / 20
    # the same step on a minibatch drawn by a permutation from a pooled rollout of 2 M rows (what train_rollout issues)
    pool = [torch.cat([t, t]) for t in d]
    idx = torch.randperm(2 * M, device=dev)[:M]
    tr.update_minibatch(*pool, idx=idx)
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        tr.update_minibatch(*pool, idx=idx)
    b.record(); torch.cuda.synchronize()
    ms_idx = a.elapsed_time(b) / 20
    # the reference's path: torch autograd + Adam on the host cores
    cpu_ms = float("nan")
    if not NOCPU:
        opt = oppo.make_optimizer(pol)
        n_cpu = 3 if M >= 16384 else 10
        oppo.minibatch_step(pol, opt, obs, act, old, adv, ret)
        t0 = time.perf_counter()
        for _ in range(n_cpu):
            oppo.minibatch_step(pol, opt, obs, act, old, adv, ret)
        cpu_ms = (time.perf_counter() - t0) / n_cpu * 1e3
    print(json.dumps(dict(kernel="dart_ppo_update", M=M, ms_flushed=round(ms, 4), ms_back_to_back=round(ms_b2b, 4), ms_permuted_minibatch=round(ms_idx, 4), launches_per_step=launches,
                          samples_per_s=M / ms_b2b * 1e3, fp32_tflops=FLOP_PER_SAMPLE * M / ms_b2b * 1e-9,
                          hbm_GBps=M * (2 * 2080 + 148) / ms_b2b * 1e-6, torch_cpu_ms=round(cpu_ms, 3), torch_cpu_threads=torch.get_num_threads(),
                          speedup_vs_torch_cpu=cpu_ms / ms_b2b)), flush=True)
    tr.close()

# rollout-time policy call (actor + critic forward + sampling) for 16 384 instances
if NOCPU:
    sys.exit(0)
B = 16384
tr = dart_b200.PPOTrainer(capacity=B)
obs, eps = torch.randn(B, 520, device=dev), torch.randn(B, 34, device=dev)
for _ in range(3):
    tr.act(obs, eps)
ts = []
for _ in range(7):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tr.act(obs, eps); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(json.dumps(dict(kernel="dart_ppo_act", B=B, ms=round(float(np.median(ts)), 4))), flush=True)

# the reference's update block at its own sizes (rollout_len 2048, mini_batch_size 64, epochs 16 -> 512 optimiser steps): eager launches
# against CUDA-graph replay of the minibatch step, and torch on the host cores
T = 2048
g = torch.Generator(device="cuda").manual_seed(0)
obs = torch.randn(T, 1, 520, device=dev, generator=g); eps = torch.randn(T, 1, 34, device=dev, generator=g)
rew = torch.randn(T, 1, device=dev, generator=g); done = torch.zeros(T, 1, device=dev)
res = {}
for mode in ("eager", "graph"):
    tr = dart_b200.PPOTrainer(capacity=2048, epochs=16, mini_batch_size=64)
    act, logp, val, _ = tr.act(obs.reshape(T, 520), eps.reshape(T, 34))
    act, logp, val = act.view(T, 1, 34), logp.view(T, 1), val.view(T, 1)
    last = val[-1].clone()
    tr.train_rollout(obs, act, logp, rew, val, done, last, generator=g, graph=(mode == "graph"))     # warm-up (+ capture)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    steps = tr.train_rollout(obs, act, logp, rew, val, done, last, generator=g, graph=(mode == "graph"))
    torch.cuda.synchronize(); res[mode] = (time.perf_counter() - t0) * 1e3
    tr.close()
pol = oppo.make_policy(3); opt = oppo.make_optimizer(pol)
o, a_, lp, ad, rt = obs.reshape(T, 520).cpu(), act.reshape(T, 34).cpu(), logp.reshape(T).cpu(), torch.randn(T), torch.randn(T)
t0 = time.perf_counter()
for e in range(2):
    perm = torch.randperm(T)
    for s in range(0, T, 64):
        i = perm[s:s + 64]
        oppo.minibatch_step(pol, opt, o[i], a_[i], lp[i], ad[i], rt[i])
cpu_ms = (time.perf_counter() - t0) * 1e3 * 8           # 2 of the 16 epochs timed
print(json.dumps(dict(path="reference-sized update block: 2048 transitions, 16 epochs x 32 minibatches of 64 = 512 optimiser steps", steps=steps,
                      eager_ms=round(res["eager"], 2), graph_ms=round(res["graph"], 2), torch_cpu_ms=round(cpu_ms, 1),
                      ms_per_step_graph=round(res["graph"] / steps, 4))), flush=True)
