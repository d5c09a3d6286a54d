"""Instance sharding across GPUs: the B200 replacement of the reference's multiprocessing fan-out.

PMPC/main_parallel.py:138-145 spawns one solver process per experiment and moves (state, target) /
(u_cmd, loss, solve_time) tuples through mp.Queue.  Here every rank (one process per GPU, torch.distributed)
owns a contiguous range of the instance axis, solves it with its own engine -- the instances are independent,
so nothing is exchanged during the solve -- and the per-instance result rows are gathered with ONE collective
(NCCL all_gather over NVLink on GPUs; gloo in the CPU tests of this host logic).
"""
import numpy as np

RESULT_COLS = 5   # u0x, u0y, J, status, iters


def shard_bounds(B, world, rank):
    """Contiguous [lo, hi) of rank's shard; the first B % world ranks take one extra instance."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_rows(out):
    """dict(u0 [n,2], J [n], status [n], iters [n]) -> float64 rows [n, 5] (numpy or torch, same backend as input)."""
    u0, J, st, it = out["u0"], out["J"], out["status"], out["iters"]
    if isinstance(u0, np.ndarray):
        return np.concatenate([u0, J[:, None], st[:, None].astype(np.float64), it[:, None].astype(np.float64)], axis=1)
    import torch
    return torch.cat([u0, J[:, None], st[:, None].to(torch.float64), it[:, None].to(torch.float64)], dim=1)


def unpack_rows(rows):
    if isinstance(rows, np.ndarray):
        return dict(u0=rows[:, :2].copy(), J=rows[:, 2].copy(), status=rows[:, 3].astype(np.int32), iters=rows[:, 4].astype(np.int32))
    import torch
    return dict(u0=rows[:, :2].contiguous(), J=rows[:, 2].contiguous(), status=rows[:, 3].to(torch.int32), iters=rows[:, 4].to(torch.int32))


class ShardedSolver:
    """Solve a global batch with every rank working on its own shard, then gather the result rows on all ranks.

    ``solve_local(x0, ref, aux) -> dict(u0, J, status, iters)`` is this rank's engine (``NMPCEngine.solve`` on host
    arrays, or a wrapper of ``solve_device`` on CUDA tensors).  Inputs may be the global arrays (every rank slices
    its own range) or, with ``local=True``, already the rank's shard.
    """

    def __init__(self, solve_local, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.solve_local = solve_local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def solve(self, B, x0, ref, aux=None, local=False):
        import torch
        lo, hi = shard_bounds(B, self.world, self.rank)
        if not local:
            x0, ref = x0[lo:hi], ref[lo:hi]
            aux = None if aux is None else aux[lo:hi]
        out = self.solve_local(x0, ref, aux)
        rows = pack_rows(out)
        if self.world == 1:
            return unpack_rows(rows)
        is_np = isinstance(rows, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(rows)) if is_np else rows.contiguous()
        # ragged shards: pad to the largest shard so one all_gather_into_tensor suffices
        nmax = (B + self.world - 1) // self.world
        pad = torch.zeros((nmax, RESULT_COLS), dtype=torch.float64, device=t.device)
        pad[: hi - lo] = t
        full = torch.empty((self.world * nmax, RESULT_COLS), dtype=torch.float64, device=t.device)
        self.dist.all_gather_into_tensor(full, pad, group=self.group)
        pieces = []
        for r in range(self.world):
            a, b = shard_bounds(B, self.world, r)
            pieces.append(full[r * nmax: r * nmax + (b - a)])
        allrows = torch.cat(pieces, dim=0)
        return unpack_rows(allrows.numpy() if is_np else allrows)
