#!/usr/bin/env python
"""Headline benchmark: converged tray-tilt NMPC solves/sec (BASELINE.json metric).

Workload (config.workload): BASELINE config 2 -- PMPC batched, the 18 shape x mass x friction objects x 64
random (x0, target) pairs = 1152 independent NLPs per step, cold-started as the reference does.  A "step"
is one pass of the hot path (one batched solve) over that batch.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

* `value`   solves/s with inputs resident in HBM (device-pointer C ABI, CUDA-event timed per step, an L2
            flush between steps outside the timed events, max over ranks).
* `e2e`     the same metric through the public host API (dart_solve_host: pinned staging, H2D, solve, D2H).
* `roofline` FP64 FMA pipe: algorithmic flops (66.9 kflop per PMPC interior-point iteration x the
            iterations actually taken) / solve-kernel time, against the DFMA peak measured in this run.
* `cpu_baseline` the oracle (oracle/ipm.py, a numpy port; NOT CasADi/IPOPT, which cannot be installed here)
            on a bounded sample of the same batch.
With --impl reference the oracle port runs through a process-per-core fan-out (the reference's
main_parallel.py launcher pattern) on the same workload; rank 0 only.
N > 1: one process per GPU (torchrun), every rank solves its own seeded batch (weak scaling), result rows
are gathered to all ranks with one NCCL all_gather per step, overlapped with the next step's solve.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PMPC_FLOPS_PER_ITER = 66.9e3      # N * (F_ric + F_dyn), n = 6, m = 2, N = 15
STATES_PER_OBJECT = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="states per object for the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU arms
def _oracle_chunk(args):
    state, target, Qp, Qv, R, mu = args
    from oracle import ipm, problems
    r = ipm.solve(problems.pmpc_problem(state, target, Qp=Qp, Qv=Qv, R=R, mu=mu))
    return int((r["status"] == 0).sum())


def cpu_oracle_rate(states_per_object, pool=None, procs=1):
    """Oracle port over the config-2 batch with `states_per_object` states per object; `pool` = process fan-out."""
    import dart_b200
    c = dart_b200.workloads.pmpc_config2(states_per_object, seed=1)
    B = c["state"].shape[0]
    keys = ("state", "target", "Qp", "Qv", "R", "mu")
    if pool is None:
        t0 = time.perf_counter()
        ok = _oracle_chunk(tuple(c[k] for k in keys))
        dt = time.perf_counter() - t0
    else:
        idx = np.array_split(np.arange(B), procs)
        chunks = [tuple(c[k][i] for k in keys) for i in idx if len(i)]
        t0 = time.perf_counter()
        ok = sum(pool.map(_oracle_chunk, chunks))
        dt = time.perf_counter() - t0
    return ok / dt, B, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"                      # one solver process per core, as main_parallel.py fans out
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    sp = max(1, args.cpu_sample, min(cores, 64) // 2)     # at least ~9 instances per worker process
    rates, times = [], []
    with mp.get_context("spawn").Pool(cores) as pool:
        budget = time.perf_counter() + 200.0
        for i in range(args.warmup + args.steps):
            r, B, dt = cpu_oracle_rate(sp, pool, cores)
            if i >= args.warmup:
                rates.append(r); times.append(dt)
            if time.perf_counter() > budget and len(rates) >= 3:
                break
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": "NMPC solves/sec", "value": v, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": len(rates), "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PMPC batched (BASELINE config 2): 18 objects x {sp} states = {18 * sp} instances/step "
                                   f"(bounded sample of the 18x{STATES_PER_OBJECT} batch)", "N": 15},
            "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"oracle/ipm.py (numpy dense interior point; CasADi/IPOPT not installable) over "
                                       f"{18 * sp} instances, one process per core (main_parallel-style fan-out)"},
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
def ensure_built():
    lib = os.path.join(ROOT, "dart-dual-arm-non-prehensile-manipulation_b200", "lib", "libdart_b200.so")
    if not os.path.exists(lib) and int(os.environ.get("LOCAL_RANK", "0")) == 0:
        import __graft_entry__ as g
        g.build_cuda()


def run_ours(args):
    import torch
    ensure_built()
    import dart_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    K, W = args.steps, max(3, args.warmup)
    c = dart_b200.workloads.pmpc_config2(STATES_PER_OBJECT, seed=1 + rank)
    B = c["state"].shape[0]
    aux_h = np.ascontiguousarray(np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], axis=1))
    eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=local)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    x0_d, tg_d, aux_d = t(c["state"]), t(c["target"]), t(aux_h)
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
    J = torch.empty((B,), dtype=torch.float64, device=dev)
    st = torch.empty((B,), dtype=torch.int32, device=dev)
    it = torch.empty((B,), dtype=torch.int32, device=dev)
    rows = torch.empty((B, 4), dtype=torch.float64, device=dev)          # [u0x, u0y, J, status] result rows
    gathered = torch.empty((world * B, 4), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    solved_evt = torch.cuda.Event()
    if world > 1:
        eng.set_result_rows(rows)        # the solve kernel writes the packed [u0x, u0y, J, status] rows itself

    def step():
        eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
        if world > 1:
            solved_evt.record()
            comm.wait_event(solved_evt)
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(gathered, rows)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # sampled from warm-up through the timed region and the e2e loop
    for _ in range(W):
        step()
        flush.zero_()
    torch.cuda.synchronize()
    peak_tf = dart_b200.measure_fp64_tflops(local) if rank == 0 else 0.0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    l0 = eng.launch_count
    conv = 0
    iters_sum = 0
    for i in range(K):
        flush.zero_()                       # L2 flush, outside the timed events
        ev0[i].record()
        step()
        ev1[i].record()
    if world > 1:
        torch.cuda.current_stream().wait_stream(comm)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count - l0
    if world > 1:
        eng.set_result_rows(None)
    ms = np.array([a.elapsed_time(b) for a, b in zip(ev0, ev1)])
    total_s = float(ms.sum() * 1e-3)
    conv = int((st == 0).sum().item())
    iters_sum = int(it.sum().item())
    tt = torch.tensor([total_s], dtype=torch.float64, device=dev)
    cv = torch.tensor([conv], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cv, op=dist.ReduceOp.SUM)
    total_s_max = float(tt.item())
    value = float(cv.item()) * K / total_s_max

    # ---- end to end through the public host API (pinned staging + H2D + solve + D2H inside the timed region)
    for _ in range(3):
        eng.solve(c["state"], c["target"], aux=aux_h, want_w=False)
    if world > 1:
        dist.barrier()
    Ke = max(10, min(K, 100))
    lat = []
    t0 = time.perf_counter()
    for _ in range(Ke):
        t1 = time.perf_counter()
        out = eng.solve(c["state"], c["target"], aux=aux_h, want_w=False)
        lat.append(time.perf_counter() - t1)
    e2e_s = time.perf_counter() - t0
    e2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
    e2e_conv = int((out["status"] == 0).sum())
    e2e_value = world * e2e_conv * Ke / float(e2.item())
    h2d = B * (6 + 6 + 4) * 8
    d2h = B * (2 + 1) * 8 + B * 2 * 4

    clk = clocks.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- single-instance latency (B = 1) through the same host API
    c1 = dart_b200.workloads.pmpc_config1()
    one = []
    for i in range(60):
        t1 = time.perf_counter()
        eng.solve(c1["state"], c1["target"], want_w=False)
        if i >= 10:
            one.append(time.perf_counter() - t1)

    # ---- throughput variant: the same distribution tiled to 2^17 instances
    big = {}
    try:
        cb = dart_b200.workloads.pmpc_config2(7282, seed=1)          # 18 * 7282 = 131076
        Bb = cb["state"].shape[0]
        xb, tb = t(cb["state"]), t(cb["target"])
        ab = t(np.stack([cb["Qp"], cb["Qv"], cb["R"], cb["mu"]], axis=1))
        ub = torch.empty((Bb, 2), dtype=torch.float64, device=dev); Jb = torch.empty((Bb,), dtype=torch.float64, device=dev)
        sb = torch.empty((Bb,), dtype=torch.int32, device=dev); ib = torch.empty((Bb,), dtype=torch.int32, device=dev)
        for _ in range(2):
            eng.solve_device(xb, tb, aux=ab, u0_out=ub, J_out=Jb, status=sb, iters=ib)
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(3):
            eng.solve_device(xb, tb, aux=ab, u0_out=ub, J_out=Jb, status=sb, iters=ib)
        b_.record()
        torch.cuda.synchronize()
        sec = a_.elapsed_time(b_) * 1e-3 / 3
        itb = int(ib.sum().item())
        big = {"instances": Bb, "solves_per_s": float((sb == 0).sum().item()) / sec, "ms": sec * 1e3,
               "fp64_tflops": itb * PMPC_FLOPS_PER_ITER / sec / 1e12, "launch": eng.last_launch_config()}
    except Exception as e:      # the variant is informative only; never fail the bench line on it
        big = {"error": str(e)[:200]}
    eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
    torch.cuda.synchronize()
    launch_cfg = eng.last_launch_config()

    # ---- roofline of the solve kernel (it is the only kernel in the step at N = 1)
    kern_s = float(np.mean(ms)) * 1e-3
    flops = iters_sum * PMPC_FLOPS_PER_ITER
    achieved = flops / kern_s / 1e12
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf if peak_tf else None,
                "traffic": 215552,   # dram__bytes_read + write per launch, ncu --set full capture (profiles/r1_summary.md)
                "peak_source": "measured in this run (dart_measure_fp64_tflops DFMA microbenchmark); MEASURED_PEAKS.json "
                               "has no FP64 entry",
                "kernel": "nmpc_solve_kernel<PmpcAxis,16,15>", "algorithmic_flops_per_launch": flops,
                "mean_iters": iters_sum / B, "note": "latency-bound: 1152 instances occupy a fraction of the SMs; see "
                                                      "throughput_variant for the filled-GPU figure"}

    cpu = None
    if not args.no_cpu_baseline:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        r, Bc, dt = cpu_oracle_rate(args.cpu_sample)
        cpu = {"value": r, "unit": "solves/s", "cores": 1, "kind": "port",
               "sample": f"oracle/ipm.py (numpy port; CasADi/IPOPT not installable) on {Bc} instances of the same batch, "
                         f"{dt:.1f} s; reference README quotes 80-100 solves/s per IPOPT worker (PMPC/README.md:266)"}

    line = {"metric": "NMPC solves/sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_s_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PMPC batched (BASELINE config 2): 18 objects x {STATES_PER_OBJECT} states = {B} "
                                   f"instances/step/GPU, cold start, tol 1e-8", "N": 15, "instances_per_gpu": B,
                       "l2": "256 MiB flush write between timed steps", "launch": launch_cfg,
                       "parallelism": f"instance sharding x{world}" + (", NCCL all_gather of result rows overlapped" if world > 1 else "")},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "p50_batch_latency_ms": float(np.median(lat) * 1e3)},
            "p50_solve_latency_ms": {"B=1 host API": float(np.median(one) * 1e3),
                                     "per-batch/B": float(np.median(lat) * 1e3 / B)},
            "gpu_launches": int(launches), "converged": conv, "clocks": clk, "roofline": roofline,
            "cpu_baseline": cpu, "throughput_variant": big}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
