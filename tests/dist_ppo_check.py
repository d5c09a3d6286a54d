"""Data-parallel PPO step on N GPUs (one process per GPU, NCCL): launched by tests/test_gpu_ppo.py::test_data_parallel_step_two_gpus as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/dist_ppo_check.py
Checks (a) the replicas stay bitwise identical over several steps, (b) one data-parallel step on two half minibatches equals the
single-GPU step on the whole minibatch up to FP32 summation order, (c) reports the step time.  Prints DIST_PPO_OK on rank 0."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_b200                      # noqa: E402
from oracle import ppo as oppo        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    pol = oppo.make_policy(3)
    M = 1024                                                   # per rank
    g = torch.Generator().manual_seed(99)
    obs = torch.randn(world * M, 520, generator=g); eps = torch.randn(world * M, 34, generator=g)
    act, logp, val, _ = oppo.act(pol, obs, eps)
    old = logp + 0.3 * torch.randn(world * M, generator=g); adv = torch.randn(world * M, generator=g); ret = val + torch.randn(world * M, generator=g)
    full = [t.to(dev) for t in (obs, act, old, adv, ret)]
    mine = [t[rank * M:(rank + 1) * M].contiguous() for t in full]
    lr = 3e-4
    tr = dart_b200.PPOTrainer(capacity=world * M, state_dict=pol.state_dict(), device=local, lr=lr)
    p0 = tr._get()[0]
    tr.update_minibatch_distributed(*mine)
    p1 = tr._get()[0]
    ok = True
    if rank == 0:                                              # (b) against the single-GPU step on the concatenated minibatch
        one = dart_b200.PPOTrainer(capacity=world * M, state_dict=pol.state_dict(), device=local, lr=lr)
        one.update_minibatch(*full)
        q1 = one._get()[0]
        err = np.abs((p1 - p0) - (q1 - p0))
        frac = float(np.mean(err <= 1e-7 + 1e-2 * lr))
        print(f"data-parallel vs single step: max err {err.max():.2e}, within 1% of lr: {frac:.4f}", flush=True)
        ok = ok and err.max() <= 1e-7 + 2 * lr and frac >= 0.98 and np.abs(p1 - p0).max() > 0.1 * lr
        one.close()
    for _ in range(3):
        tr.update_minibatch_distributed(*mine)
    p = torch.from_numpy(tr._get()[0]).to(dev)                 # (a) replicas identical, bit for bit
    allp = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(allp, p)
    same = all(torch.equal(allp[0], q) for q in allp)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        tr.update_minibatch_distributed(*mine)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for _ in range(20):
        tr.update_minibatch(*mine)
    torch.cuda.synchronize(); ms_local = (time.perf_counter() - t0) / 20 * 1e3
    if rank == 0:
        print(f"world {world}: replicas identical {same}; data-parallel step {ms:.3f} ms ({world * M} samples), local step {ms_local:.3f} ms ({M} samples)", flush=True)
        if ok and same:
            print("DIST_PPO_OK", flush=True)
    tr.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
