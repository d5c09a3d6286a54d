"""N > 1 host logic on the CPU: world_size-2 gloo, contiguous instance sharding + one all_gather of result rows.

The local solver is the host-emulation TEST harness (the CUDA solver source compiled for the CPU); the product path
uses NMPCEngine on each rank's GPU with the same ShardedSolver.
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dart_b200
    from tests import helpers
    he = helpers.HostEmu()
    c, aux, _ = helpers.pmpc_case(1)
    cfg = dart_b200.pmpc_cfg()
    sh = dart_b200.ShardedSolver(lambda x0, ref, a: he.solve(cfg, x0, ref, a))
    out = sh.solve(B, c["state"][:B], c["target"][:B], aux[:B])
    lo, hi = dart_b200.shard_bounds(B, world, rank)
    q.put((rank, lo, hi, out["u0"], out["J"], out["status"], out["iters"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [18, 17, 1])
def test_two_rank_sharding_equals_single_rank(hostemu, B):
    import torch.multiprocessing as mp
    import dart_b200
    from tests import helpers
    c, aux, _ = helpers.pmpc_case(1)
    single = hostemu.solve(dart_b200.pmpc_cfg(), c["state"][:B], c["target"][:B], aux[:B])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = sorted((lo, hi) for _, lo, hi, *_ in res)
    assert covered[0][0] == 0 and covered[-1][1] == B and covered[0][1] == covered[1][0]
    for _, _, _, u0, J, st, it in res:          # every rank holds the full, bitwise-identical result
        assert np.array_equal(u0, single["u0"]) and np.array_equal(J, single["J"])
        assert np.array_equal(st, single["status"]) and np.array_equal(it, single["iters"])


def test_shard_bounds_cover_and_balance():
    import dart_b200
    for B in (0, 1, 7, 1152, 2 ** 20 + 3):
        for world in (1, 2, 4, 8):
            b = [dart_b200.shard_bounds(B, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == B
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
