"""Peer-memory gather of the result rows on N GPUs (one process per GPU): launched by
tests/test_gpu_nmpc.py::test_peer_rows_gather_two_gpus as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/dist_peer_rows_check.py
For PMPC, RMPC and LMPC (the three epilogues that store rows) and several steps with changing inputs: the rows the solve kernels
stored into every rank's gathered buffer over NVLink equal the NCCL all_gather of the local rows, bit for bit, after the flag
hand-shake; also times both gathers.  Prints PEER_ROWS_OK on rank 0 (PEER_ROWS_UNAVAILABLE when the node cannot do peer access)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dart_b200                      # noqa: E402

W = dart_b200.workloads


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    report = []
    for method in ("pmpc", "rmpc", "lmpc"):
        if method == "pmpc":
            c = W.pmpc_config2(8, seed=3 + rank)
            aux = np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], axis=1)
            x0, ref, cfg = c["state"], c["target"], dart_b200.pmpc_cfg()
        elif method == "rmpc":
            d = W.rmpc_inputs(96, seed=5 + rank)
            x0, ref, aux, cfg = d["x0"], d["ref"], d["aux"], dart_b200.rmpc_cfg()
        else:
            d = W.lmpc_inputs(160, seed=7 + rank)
            x0, ref, aux, cfg = d["x0"], d["ref"], d["aux"], dart_b200.lmpc_cfg()
        B = x0.shape[0]
        eng = dart_b200.NMPCEngine(cfg, device=local)
        rows = torch.zeros((B, 4), dtype=torch.float64, device=dev)
        gathered = torch.zeros((world * B, 4), dtype=torch.float64, device=dev)
        eng.set_result_rows(rows)
        peer = dart_b200.parallel.PeerRows.create(eng, B, local)
        if peer is None:
            if rank == 0:
                print("PEER_ROWS_UNAVAILABLE")
            dist.destroy_process_group()
            return
        X, R, A = t(x0), t(ref), t(aux)
        for step in range(4):
            Xs = X * (1.0 + 0.01 * step)                 # new inputs every step: stale rows would show
            out = eng.solve_device(Xs, R, aux=A)
            peer.handshake()
            dist.all_gather_into_tensor(gathered, rows)
            torch.cuda.synchronize()
            assert int(peer.timed_out.item()) == 0, "hand-shake timed out"
            if not torch.equal(peer.gathered, gathered):
                bad = (peer.gathered != gathered).nonzero()
                raise AssertionError((method, step, rank, bad.shape[0], bad[:6].tolist(), peer.gathered[bad[0, 0]].tolist(),
                                      gathered[bad[0, 0]].tolist()))
            assert torch.equal(rows[:, :2], out["u0"]) and int((gathered[:, 3] == 0).sum()) > 0
        # timing: solve + gather, both ways
        ms = {}
        for kind in ("peer", "nccl"):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                eng.solve_device(X, R, aux=A)
                if kind == "peer":
                    peer.handshake()
                else:
                    dist.all_gather_into_tensor(gathered, rows)
            e1.record(); torch.cuda.synchronize()
            ms[kind] = e0.elapsed_time(e1) / 20
        report.append((method, B, round(ms["peer"], 4), round(ms["nccl"], 4)))
        peer.close()
        eng.set_result_rows(None)
        eng.close()
    if rank == 0:
        print("PEER_ROWS_OK", report)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
