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


class _DevView:
    """A device buffer this package allocated, as something torch.as_tensor can wrap (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerRows:
    """The gather of the result rows WITHOUT a collective (GPUs of one node, one process per GPU, at most 8).

    Every rank owns a gathered buffer ``[world * rows_per_rank, 4]`` and a flag array (``dart_peer_alloc``: dedicated device
    allocations exported with CUDA IPC); every other rank maps them for its own device (``dart_peer_open``, lazy peer access)
    and registers the pointers with its engine (``dart_set_result_rows_peers``): the solve kernel then stores each instance's
    ``[u0x, u0y, J, status]`` row into ALL ranks' buffers over NVLink from its epilogue, and ``handshake()`` (one tiny launch:
    flag stores to the peers, then a wait on the own flags) tells the stream that every peer's rows of the step have landed.
    Replaces ``all_gather_into_tensor`` -- 59 us at 8 GPUs against a 80 us solve -- on the weak-scaled headline step.

    ``PeerRows.create`` returns None when the node cannot do it (no peer access, IPC refused): the caller keeps the NCCL
    gather.  Collective: every rank of the group must call it."""

    def __init__(self):
        self.step_no = 0
        self._opened, self._own = [], []

    @classmethod
    def create(cls, engine, rows_per_rank, device, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        self = cls()
        self.engine, self.L, self.C = engine, L, C
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows_per_rank, self.device = int(rows_per_rank), int(device)
        dev = torch.device("cuda", device)
        VP = C.c_void_p
        nrow = self.world * self.rows_per_rank
        ok = hasattr(L, "dart_peer_alloc") and self.world <= 8
        mine = None
        with torch.cuda.device(dev):
            if ok:
                gp, fp = VP(), VP()
                gh, fh = C.create_string_buffer(64), C.create_string_buffer(64)
                if L.dart_peer_alloc(nrow * 32, C.byref(gp), gh) == 0:
                    self._own.append(gp)
                    if L.dart_peer_alloc(8 * 8 + 64, C.byref(fp), fh) == 0:
                        self._own.append(fp)
                        mine = (gh.raw, fh.raw)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            ok = all(e is not None for e in everyone)
            rows_ptrs, flag_ptrs = [None] * self.world, [None] * self.world
            if ok:
                for r, (gh_r, fh_r) in enumerate(everyone):
                    if r == self.rank:
                        rows_ptrs[r], flag_ptrs[r] = self._own[0], self._own[1]
                        continue
                    g, f = VP(), VP()
                    if L.dart_peer_open(gh_r, C.byref(g)) != 0:
                        ok = False
                        break
                    self._opened.append(g)
                    if L.dart_peer_open(fh_r, C.byref(f)) != 0:
                        ok = False
                        break
                    self._opened.append(f)
                    rows_ptrs[r], flag_ptrs[r] = g, f
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self._release(dist, group)
                return None
            self.gathered = torch.as_tensor(_DevView(self._own[0].value, (nrow, 4), "<f8"), device=dev)
            self.timed_out = torch.zeros((1,), dtype=torch.int32, device=dev)
            self._rows_ptrs = (VP * self.world)(*rows_ptrs)
            self._flag_ptrs = (VP * self.world)(*flag_ptrs)
            _lib.check(L.dart_set_result_rows_peers(engine._h, self._rows_ptrs, self.world, self.rank * self.rows_per_rank),
                       "dart_set_result_rows_peers")
        self._dist, self._group = dist, group
        return self

    def handshake(self):
        """Call after the solve of a step, on the same stream: afterwards `gathered` holds every rank's rows of that step."""
        import torch
        from . import _lib
        self.step_no += 1
        stream = self.C.c_void_p(torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream)
        _lib.check(self.L.dart_peer_handshake(self._flag_ptrs, self.world, self.rank, self.step_no,
                                              self.C.c_void_p(self.timed_out.data_ptr()), stream), "dart_peer_handshake")

    def _release(self, dist, group):
        import torch
        torch.cuda.synchronize(torch.device("cuda", self.device))
        for p in self._opened:
            self.L.dart_peer_close(p)
        self._opened = []
        dist.barrier(group=group)                  # nobody frees a buffer a peer still maps
        for p in self._own:
            self.L.dart_peer_free(p)
        self._own = []

    def close(self):
        """Collective: detach from the engine, unmap the peers' buffers, free the own ones."""
        self.L.dart_set_result_rows_peers(self.engine._h, None, 0, 0)
        self.gathered = None
        self._release(self._dist, self._group)
