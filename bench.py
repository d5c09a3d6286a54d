#!/usr/bin/env python
"""Headline benchmark: converged tray-tilt NMPC solves/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (config.workload): BASELINE config 2 -- PMPC batched, the 18 shape x mass x friction objects x 64
random (x0, target) pairs = 1152 independent NLPs per step and GPU, cold-started as the reference does.  A "step" is
one pass of the hot path (one batched solve) over that batch.

* `value`    solves/s with inputs resident in HBM (device-pointer C ABI, CUDA-event timed per step, an L2 flush between
             steps outside the timed events, max over ranks).  N > 1: every rank solves its own seeded batch (weak
             scaling) and ONE NCCL all_gather of the result rows per step runs INSIDE the timed events.
* `e2e`      the same metric through the public host API (dart_solve_host: host arrays -> pinned staging -> device over the
             host link (kernel-side reads of the mapped block below 1 MB, copy engines above) -> solve -> results back); N > 1:
             pinned H2D + dart_solve + all_gather + D2H of the gathered rows.
* `roofline` FP64 FMA pipe: algorithmic flops (66.9 kflop per PMPC interior-point iteration + 4.7 kflop for the
             corrector's second solve, x the iterations actually taken) / solve-kernel time, against the DFMA peak measured
             in this run.  The predictor-corrector steps cut the iterations (9.0 -> 5.7 on average, 15 -> 10 for the slowest
             instance), so solves/s rises while this fraction, which credits flops and not solves, does not.
* `cpu_baseline` the oracle (oracle/ipm.py, a numpy port; NOT CasADi/IPOPT, which cannot be installed here) on a bounded
             sample of the same batch, one core.
* `scale_sweep` (every N) BASELINE config 5: 2^20 mixed PMPC/RMPC/LMPC instances (1/3 each, re-seeded per shard),
             contiguous shards over the N ranks -- STRONG scaling --, one all_gather of [u0, J, status] rows per pass,
             the timed region (>= 1 s) ends after the gather; reports solves/s, gather_ms, per-method iterations,
             converged count and the per-rank times.
* `configs`  (N = 1) driver-visible sub-records of BASELINE configs 3 (RMPC + RLS closed loop, 4096 x 256 steps) and
             4 (LMPC + policy MLP closed loop, 16384 x 64 steps), each with its kernel's roofline fraction and converged count.

--impl reference: the CPU arm.  The oracle port behind the reference launcher's protocol (PMPC/main_parallel.py:10-43:
one solver process per core, `(state, target)` items in, `(u_cmd, loss, solve_time)` replies out, "STOP" to end) on the
FULL 1152-instance batch of the GPU arm; rank 0 only.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# N * (F_ric + F_dyn) per IPM iteration.  Cold-started launches run predictor-corrector iterations (DESIGN.md
# section 2; the warm-started RMPC closed loop of config 3 stays on the monotone schedule): one factorisation and TWO solves, so the dense count gains
# the vector parts of a second sweep pair, N * [(4 n^2 + 4 n m) + (2 n (n + m) + 2 n m)] = 15 * 312 = 4.7 k at n = 6, m = 2
# (PMPC; the monotone method's figure is 66.9 k) and 20 * 760 = 15.2 k at n = 10, m = 2 (LMPC; monotone 245.7 k).
_FLOPS_MONOTONE = {"pmpc": 66.9e3, "rmpc": 55.3e3, "lmpc": 245.7e3}
_FLOPS_CORRECTOR = {"pmpc": 4.68e3, "rmpc": 20 * 312.0, "lmpc": 15.2e3}
_PC = {"pmpc": True, "rmpc": True, "lmpc": True}             # cold-started launches under DART_BARRIER_AUTO (csrc/models.cuh)
if os.environ.get("DART_BARRIER_MONOTONE"):
    _PC = {k: False for k in _PC}
elif os.environ.get("DART_BARRIER_MEHROTRA"):
    _PC = {k: True for k in _PC}
FLOPS_PER_ITER = {k: _FLOPS_MONOTONE[k] + (_FLOPS_CORRECTOR[k] if _PC[k] else 0.0) for k in _PC}
PMPC_FLOPS_PER_ITER = FLOPS_PER_ITER["pmpc"]
STATES_PER_OBJECT = 64
SWEEP_TOTAL = 2 ** 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="states per object for the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-5 scale sweep sub-record")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 3/4 sub-records")
    ap.add_argument("--sweep-total", type=int, default=SWEEP_TOTAL)
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


def cpu_oracle_rate(states_per_object):
    """Oracle port over a config-2 sample with `states_per_object` states per object, this process, one batch."""
    import dart_b200
    c = dart_b200.workloads.pmpc_config2(states_per_object, seed=1)
    B = c["state"].shape[0]
    t0 = time.perf_counter()
    ok = _oracle_chunk(tuple(c[k] for k in ("state", "target", "Qp", "Qv", "R", "mu")))
    dt = time.perf_counter() - t0
    return ok / dt, B, dt


def oracle_mpc_worker(objects, state_queue, control_queue):
    """The CPU arm's solver process.  Same service loop as the reference's ``mpc_worker`` (PMPC/main_parallel.py:10-43,
    mpc_3d.py:140-158): build the controller(s), then `(state, target)` items in, `(u_cmd, loss, solve_time)` replies out,
    "STOP" ends it.  One difference forced by 18 objects sharing fewer cores: an item carries the index of its object
    (the reference starts one process per experiment, each with one object's parameters)."""
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ipm, problems
    while True:
        item = state_queue.get()
        if isinstance(item, str) and item == "STOP":
            break
        state, target, obj = item
        o = objects[obj]
        t0 = time.perf_counter()
        r = ipm.solve(problems.pmpc_problem(np.asarray(state)[None], np.asarray(target)[None], Qp=o["Qp"], Qv=o["Qv"], R=o["R"], mu=o["mu"]))
        solve_time = time.perf_counter() - t0
        control_queue.put((r["U"][0, 0].copy(), np.array([r["J"][0]]), solve_time, int(r["status"][0])))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"                      # one solver process per core, as main_parallel.py fans out
    import multiprocessing as mp
    import dart_b200
    cores = os.cpu_count() or 1
    c = dart_b200.workloads.pmpc_config2(STATES_PER_OBJECT, seed=1)          # the GPU arm's batch (rank 0)
    B = c["state"].shape[0]
    objects = dart_b200.workloads.pmpc_objects()
    obj_of = np.repeat(np.arange(len(objects)), STATES_PER_OBJECT)
    ctx = mp.get_context("spawn")                                             # mpc_3d.py:161
    workers = []
    for _ in range(cores):
        sq, cq = ctx.Queue(), ctx.Queue()
        p = ctx.Process(target=oracle_mpc_worker, args=(objects, sq, cq), daemon=True)
        p.start()
        workers.append((p, sq, cq))

    def one_pass():
        t0 = time.perf_counter()
        counts = [0] * cores
        for i in range(B):
            w = i % cores
            workers[w][1].put((c["state"][i], c["target"][i], int(obj_of[i])))
            counts[w] += 1
        ok, st = 0, []
        for w in range(cores):
            for _ in range(counts[w]):
                u_cmd, loss, solve_time, status = workers[w][2].get()
                ok += status == 0
                st.append(solve_time)
        return ok, time.perf_counter() - t0, st

    one_pass()                                   # first pass pays the workers' imports; never timed
    rates, times, solve_times, conv = [], [], [], 0
    budget = time.perf_counter() + 240.0
    W = max(0, args.warmup - 1)
    for i in range(W + args.steps):
        ok, dt, st = one_pass()
        if i >= W:
            rates.append(ok / dt); times.append(dt); solve_times += st; conv = ok
        if time.perf_counter() > budget and len(rates) >= 3:
            break
    for p, sq, cq in workers:
        sq.put("STOP")
    for p, sq, cq in workers:
        p.join(5)
    v = float(np.mean(rates))
    stms = np.array(solve_times) * 1e3
    line = {"impl": "reference", "metric": "NMPC solves/sec", "value": v, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": len(rates), "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PMPC batched (BASELINE config 2): 18 objects x {STATES_PER_OBJECT} states = {B} "
                                   f"instances/step, cold start, tol 1e-8", "N": 15, "instances_per_gpu": B},
            "converged": int(conv),
            "p50_solve_latency_ms": float(np.median(stms)), "p99_solve_latency_ms": float(np.percentile(stms, 99)),
            "cpu_baseline": {"value": v, "unit": "solves/s", "cores": cores, "kind": "port",
                             "sample": f"the full {B}-instance batch per step: oracle/ipm.py (numpy dense interior point; CasADi/IPOPT "
                                       f"not installable) in {cores} solver processes (os.cpu_count() = {cores}) behind the reference "
                                       f"launcher's (state, target) -> (u_cmd, loss, solve_time) queue protocol, one instance per item; "
                                       f"reference README quotes 80-100 solves/s per IPOPT worker (PMPC/README.md:266)"},
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
def ensure_built():
    lib = os.path.join(ROOT, "dart-dual-arm-non-prehensile-manipulation_b200", "lib", "libdart_b200.so")
    if not os.path.exists(lib) and int(os.environ.get("LOCAL_RANK", "0")) == 0:
        import __graft_entry__ as g
        g.build_cuda()


def _ev(torch):
    return torch.cuda.Event(enable_timing=True)


def scale_sweep(torch, dist, dart_b200, dev, local, world, rank, total, peak_tf):
    """BASELINE config 5: `total` mixed instances, 1/3 per method, every rank a contiguous shard of each
    method (methods interleaved across ranks, not blocked by method), inputs re-seeded per shard (seed = 100 + rank),
    one all_gather of [u0x, u0y, J, status] rows per pass.  Strong scaling: `total` is fixed as N grows."""
    W = dart_b200.workloads
    third = total // 3
    lo, hi = dart_b200.shard_bounds(third, world, rank)
    n = hi - lo
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    seed = 100 + rank
    S = (n + 17) // 18
    cp = W.pmpc_config2(S, seed=seed)
    ins = [(t(cp["state"][:n]), t(cp["target"][:n]), t(np.stack([cp["Qp"], cp["Qv"], cp["R"], cp["mu"]], 1)[:n]))]
    rd = W.rmpc_inputs(n, seed=seed)
    ins.append((t(rd["x0"]), t(rd["ref"]), t(rd["aux"])))
    ld = W.lmpc_inputs(n, seed=seed)
    ins.append((t(ld["x0"]), t(ld["ref"]), t(ld["aux"])))
    names = ("pmpc", "rmpc", "lmpc")
    engs = [dart_b200.NMPCEngine(f(), device=local) for f in (dart_b200.pmpc_cfg, dart_b200.rmpc_cfg, dart_b200.lmpc_cfg)]
    nmax = (third + world - 1) // world
    # the solve kernels write their packed result rows straight into this rank's slice of the gather buffer
    pad = torch.zeros((3 * nmax, 4), dtype=torch.float64, device=dev)
    for i, e in enumerate(engs):
        e.set_result_rows(pad[i * nmax: i * nmax + n])
    full = torch.empty((world * 3 * nmax, 4), dtype=torch.float64, device=dev) if world > 1 else None
    outs = [dict(u0_out=torch.empty((n, 2), dtype=torch.float64, device=dev), J_out=torch.empty((n,), dtype=torch.float64, device=dev),
                 status=torch.empty((n,), dtype=torch.int32, device=dev), iters=torch.empty((n,), dtype=torch.int32, device=dev)) for _ in range(3)]

    def one_pass():
        for e, (x, r, a), o in zip(engs, ins, outs):
            e.solve_device(x, r, aux=a, **o)
        if world > 1:
            torch.cuda.nvtx.range_push("gather")
            dist.all_gather_into_tensor(full, pad)        # synchronous op: the current stream waits for the collective
            torch.cuda.nvtx.range_pop()

    one_pass(); torch.cuda.synchronize()
    a, b = _ev(torch), _ev(torch)
    a.record(); one_pass(); b.record(); torch.cuda.synchronize()
    tw = torch.tensor([a.elapsed_time(b) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    passes = max(3, int(math.ceil(1.05 / float(tw.item()))))             # timed region >= 1 s
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = _ev(torch), _ev(torch)
    solve_evs = []
    e0.record()
    for _ in range(passes):
        one_pass()
    e1.record()                                   # after the last gather has been waited for on this stream
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sec_local = e0.elapsed_time(e1) * 1e-3
    # the gather alone (same buffers), and the solves alone per method, outside the headline clock
    gather_ms = 0.0
    if world > 1:
        g0, g1 = _ev(torch), _ev(torch)
        dist.barrier(); torch.cuda.synchronize()
        g0.record()
        for _ in range(5):
            dist.all_gather_into_tensor(full, pad)
        g1.record(); torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1) / 5
    per_method_ms = []
    for e, (x, r, aa), o in zip(engs, ins, outs):
        m0, m1 = _ev(torch), _ev(torch)
        m0.record(); e.solve_device(x, r, aux=aa, **o); m1.record(); torch.cuda.synchronize()
        per_method_ms.append(m0.elapsed_time(m1))
    for e in engs:
        e.set_result_rows(None)
    conv_local = float(sum((o["status"] == 0).sum().item() for o in outs))
    iters_local = [float(o["iters"].double().sum().item()) for o in outs]
    stats = torch.tensor([sec_local, conv_local, gather_ms] + iters_local + per_method_ms + [float(n)], dtype=torch.float64, device=dev)
    allst = stats[None]
    if world > 1:
        buf = torch.empty((world, stats.numel()), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(buf, stats)
        allst = buf
    A = allst.cpu().numpy()
    sec = float(A[:, 0].max())
    conv = float(A[:, 1].sum())
    iters_tot = A[:, 3:6].sum(axis=0)
    nsum = float(A[:, 9].sum())
    flops = sum(iters_tot[i] * FLOPS_PER_ITER[names[i]] for i in range(3)) * passes
    ach = flops / sec / 1e12
    rec = {"workload": f"BASELINE config 5: {3 * third} mixed instances (1/3 PMPC, RMPC, LMPC), contiguous shards over {world} GPU(s), "
                       f"one all_gather of [u0, J, status] rows per pass inside the timed region",
           "scaling": "strong", "total_instances": 3 * third, "n_gpus": world, "passes": passes, "seconds": sec,
           "solves_per_s": conv * passes / sec, "converged": int(conv), "ms_per_pass": sec / passes * 1e3,
           "gather_ms": float(A[:, 2].max()), "gather_bytes": int(world * 3 * nmax * 32) if world > 1 else 0,
           "per_method_mean_iters": {names[i]: float(iters_tot[i] / nsum) for i in range(3)},
           "per_method_solve_ms_rank_max": {names[i]: float(A[:, 6 + i].max()) for i in range(3)},
           "per_rank_seconds": [float(v) for v in A[:, 0]],
           "roofline": {"bound": "fp64", "achieved": ach, "peak": peak_tf * world if peak_tf else None, "unit": "TFLOP/s",
                        "frac": ach / (peak_tf * world) if peak_tf else None},
           "limiting": "per-rank solve time (sum of the three method kernels; LMPC dominates); the gather is "
                       f"{float(A[:, 2].max()):.2f} ms of {sec / passes * 1e3:.1f} ms per pass"}
    for e in engs:
        e.close()
    return rec


def config3_record(torch, dart_b200, dev, local, peak_tf, B=4096, T=256):
    """BASELINE config 3: RMPC + per-instance RLS closed loop on the surrogate plant."""
    c = dart_b200.workloads.rmpc_config3(B, seed=2)
    x = torch.from_numpy(c["x0"]).to(dev)
    ctl = dart_b200.RMPCBatch(B, c["target"], c["x0"], device=local)
    rv0 = np.zeros((B, 4)); rv0[:, [0, 2]] = c["x0"][:, [0, 2]]
    ctl.set_virtual_reference(rv0)
    mu = torch.from_numpy(c["mu_plant"]).to(dev); cp = torch.from_numpy(c["c_plant"]).to(dev)
    stat = torch.zeros(5, dtype=torch.int64, device=dev); it_sum = torch.zeros((), dtype=torch.int64, device=dev)
    a, b = _ev(torch), _ev(torch)
    a.record()
    for _ in range(T):
        u = ctl.step(x)
        stat += torch.bincount(ctl.status.long(), minlength=5)
        it_sum += ctl.iters.sum()
        x = dart_b200.rmpc_plant_step_device(x, u, mu, cp)
    b.record(); torch.cuda.synchronize()
    sec = a.elapsed_time(b) * 1e-3
    st = stat.cpu().numpy()
    its = float(it_sum.item())
    ach = its * _FLOPS_MONOTONE["rmpc"] / sec / 1e12          # warm-started solves: monotone schedule
    # the solve kernel alone on mid-episode inputs (one launch; the closed loop above also runs the RLS prologue and the plant)
    d = dart_b200.workloads.rmpc_inputs(B)
    k = _time_kernel(torch, dart_b200, dev, local, "rmpc", dart_b200.rmpc_cfg(), d["x0"], d["ref"], d["aux"], peak_tf)
    ctl.engine.close()
    return {"workload": f"BASELINE config 3: RMPC + RLS closed loop, {B} instances x {T} steps, surrogate plant", "B": B, "steps": T,
            "seconds": sec, "solves_per_s": float(st[0]) / sec, "converged": int(st[0]), "solves": B * T,
            "status_counts": {"converged": int(st[0]), "max_iter": int(st[1]), "infeasible": int(st[2]), "numeric": int(st[3])},
            "mean_iters": its / (B * T), "roofline": {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                                                     "frac": ach / peak_tf if peak_tf else None}, "kernel": k}


def config4_record(torch, dart_b200, dev, local, peak_tf, B=16384, T=64):
    """BASELINE config 4: LMPC with the policy MLP (Policy(520,34,{}) under seed 3), model-as-plant closed loop."""
    c = dart_b200.workloads.lmpc_config4(B, seed=3)
    ctl = dart_b200.LMPCBatch(B, c["pvec"], seed=3, device=local)
    x = torch.from_numpy(c["state"]).to(dev); tg = torch.from_numpy(c["target"]).to(dev)
    stat = torch.zeros(5, dtype=torch.int64, device=dev); it_sum = torch.zeros((), dtype=torch.int64, device=dev)
    a, b = _ev(torch), _ev(torch)
    a.record()
    for _ in range(T):
        ctl.step(x, tg)
        stat += torch.bincount(ctl.status.long(), minlength=5)
        it_sum += ctl.iters.sum()
        x = ctl.w[:, 8:16].contiguous()          # plant = the controller's own model: predicted x_1 of the optimal plan
    b.record(); torch.cuda.synchronize()
    sec = a.elapsed_time(b) * 1e-3
    st = stat.cpu().numpy()
    its = float(it_sum.item())
    ach = its * FLOPS_PER_ITER["lmpc"] / sec / 1e12
    d = dart_b200.workloads.lmpc_inputs(B)
    k = _time_kernel(torch, dart_b200, dev, local, "lmpc", dart_b200.lmpc_cfg(), d["x0"], d["ref"], d["aux"], peak_tf)
    mlp = _time_mlp(torch, dart_b200, dev, local, B)
    ctl.engine.close(); ctl.policy.close()
    return {"workload": f"BASELINE config 4: LMPC + policy MLP closed loop, {B} instances x {T} steps, model-as-plant", "B": B, "steps": T,
            "seconds": sec, "solves_per_s": float(st[0] + st[4]) / sec, "converged": int(st[0] + st[4]), "solves": B * T,
            "status_counts": {"converged": int(st[0]), "max_iter": int(st[1]), "infeasible": int(st[2]), "numeric": int(st[3]), "acceptable": int(st[4])},
            "mean_iters": its / (B * T), "roofline": {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                                                     "frac": ach / peak_tf if peak_tf else None}, "kernel": k, "policy_mlp": mlp}


def _time_kernel(torch, dart_b200, dev, local, method, cfg, x0, ref, aux, peak_tf, reps=3):
    eng = dart_b200.NMPCEngine(cfg, local)
    B = x0.shape[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    u0 = torch.empty((B, 2), dtype=torch.float64, device=dev); J = torch.empty((B,), dtype=torch.float64, device=dev)
    st = torch.empty((B,), dtype=torch.int32, device=dev); it = torch.empty((B,), dtype=torch.int32, device=dev)
    X, R, A = t(x0), t(ref), t(aux)
    for _ in range(2):
        eng.solve_device(X, R, aux=A, u0_out=u0, J_out=J, status=st, iters=it)
    torch.cuda.synchronize()
    a, b = _ev(torch), _ev(torch)
    a.record()
    for _ in range(reps):
        eng.solve_device(X, R, aux=A, u0_out=u0, J_out=J, status=st, iters=it)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    its = int(it.sum().item()); ok = int((st == 0).sum().item())
    ach = its * FLOPS_PER_ITER[method] / (ms * 1e-3) / 1e12
    rec = {"what": f"{method} solve kernel alone, one launch over {B} mid-episode instances (cold start)", "ms": ms,
           "solves_per_s": ok / ms * 1e3, "converged": ok, "mean_iters": its / B,
           "roofline": {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf if peak_tf else None},
           "launch": eng.last_launch_config()}
    eng.close()
    return rec


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json (driver-measured copy bandwidth)"
    except Exception:
        return 6552.0, "fallback 6552 GB/s (B200_PROFILING.md)"


def _time_mlp(torch, dart_b200, dev, local, B):
    hbm, src = _hbm_peak()
    obs = torch.randn((B, 520), dtype=torch.float32, device=dev); out = torch.empty((B, 34), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rec = {}
    for prec in ("fp32", "tf32"):
        pol = dart_b200.PolicyMLP(seed=3, device=local, precision=prec)
        for _ in range(3):
            pol.forward(obs, out)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = _ev(torch), _ev(torch)
            a.record(); pol.forward(obs, out); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        gbs = B * (2080 + 136) / (ms * 1e-3) / 1e9
        pol.close()
        rec[prec] = {"ms": ms, "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}}
    return {"what": f"policy MLP forward, B = {B}, L2 flushed between launches; 'fp32' = the default kernel (3xTF32 layer 1 on "
                    f"tcgen05 + FP32 layers 2-3, <= 2e-5 of the reference's FP32 forward), 'tf32' = single-pass TF32 (<= 8e-3)",
            "algorithmic_bytes_per_instance": 2216, "peak_source": src, **rec}


def _ncu_traffic():
    """dram bytes per launch of the headline kernel, from the committed ncu export (profiles/r2_pmpc_ncu_traffic.json,
    written by tools/ncu_traffic.py from an `ncu --set full` capture of this command); None when absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_pmpc_ncu_traffic.json")))
        return int(d["dram_bytes_read"] + d["dram_bytes_write"]), d.get("source")
    except Exception:
        return None, None


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
    peer = None
    gather_kind = "none"
    if world > 1:
        eng.set_result_rows(rows)        # the solve kernel writes the packed [u0x, u0y, J, status] rows itself
        gather_kind = "NCCL all_gather_into_tensor of the result rows"
        if not os.environ.get("DART_BENCH_NCCL_GATHER"):
            # gather without a collective: the solve kernel stores the rows into every rank's buffer over NVLink (peer memory),
            # one hand-shake launch per step; checked here against the NCCL gather before it is used
            try:
                peer = dart_b200.parallel.PeerRows.create(eng, B, local)
            except Exception as e:
                peer = None
                if rank == 0:
                    print(f"peer-rows gather unavailable ({e!r}); using NCCL", file=sys.stderr)
            if peer is not None:
                eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
                peer.handshake()
                dist.all_gather_into_tensor(gathered, rows)
                torch.cuda.synchronize()
                same = torch.tensor([1 if (torch.equal(peer.gathered, gathered) and int(peer.timed_out.item()) == 0) else 0],
                                    dtype=torch.int32, device=dev)
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
                if int(same.item()) == 0:
                    peer.close()
                    peer = None
                    if rank == 0:
                        print("peer-rows gather did not reproduce the NCCL gather; using NCCL", file=sys.stderr)
            if peer is not None:
                gather_kind = ("peer memory: the solve kernel stores every row into all ranks' gathered buffers over NVLink "
                               "(CUDA IPC), one flag hand-shake launch per step; verified equal to the NCCL all_gather at start-up")

    def step():
        eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
        if world > 1:
            torch.cuda.nvtx.range_push("gather")
            if peer is not None:
                peer.handshake()                             # the rows are already on their way: wait for every peer's flag
            else:
                dist.all_gather_into_tensor(gathered, rows)  # synchronous: the launching stream waits for the collective
            torch.cuda.nvtx.range_pop()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # sampled from warm-up through the timed region and the e2e loop
    for _ in range(W):
        step()
        flush.zero_()
    torch.cuda.synchronize()
    peak_tf = dart_b200.measure_fp64_tflops(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0 = [_ev(torch) for _ in range(K)]
    ev1 = [_ev(torch) for _ in range(K)]
    l0 = eng.launch_count
    for i in range(K):
        flush.zero_()                       # L2 flush, outside the timed events
        ev0[i].record()
        step()
        ev1[i].record()                     # N > 1: recorded after the stream has waited for the all_gather
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count - l0 + (K if peer is not None else 0)      # + one hand-shake launch per step
    ms = np.array([a.elapsed_time(b) for a, b in zip(ev0, ev1)])
    total_s = float(ms.sum() * 1e-3)
    conv = int((st == 0).sum().item())
    iters_sum = int(it.sum().item())
    # kernel-only time of the solve (for the roofline) when the step also holds the gather
    kern_ms = float(np.mean(ms))
    gather_ms = 0.0
    nccl_gather_ms = 0.0
    if world > 1:
        k0, k1 = _ev(torch), _ev(torch)
        torch.cuda.synchronize()
        k0.record()
        for _ in range(10):
            eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
        k1.record(); torch.cuda.synchronize()
        kern_ms = k0.elapsed_time(k1) / 10
        dist.barrier()
        g0, g1 = _ev(torch), _ev(torch)
        g0.record()
        for _ in range(10):
            if peer is not None:
                peer.handshake()
            else:
                dist.all_gather_into_tensor(gathered, rows)
        g1.record(); torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1) / 10
        g0.record()
        for _ in range(10):
            dist.all_gather_into_tensor(gathered, rows)
        g1.record(); torch.cuda.synchronize()
        nccl_gather_ms = g0.elapsed_time(g1) / 10
        eng.set_result_rows(None)
    tt = torch.tensor([total_s], dtype=torch.float64, device=dev)
    cv = torch.tensor([conv], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cv, op=dist.ReduceOp.SUM)
    total_s_max = float(tt.item())
    value = float(cv.item()) * K / total_s_max

    # ---- end to end (host buffers in, host results out, copies inside the timed region)
    Ke = max(10, min(K, 100))
    lat = []
    if world == 1:
        out = eng.make_out(B, want_w=False)           # result arrays re-used by every call (numpy's out= idiom)
        for _ in range(3):
            eng.solve(c["state"], c["target"], aux=aux_h, want_w=False, out=out)
        t0 = time.perf_counter()
        for _ in range(Ke):
            t1 = time.perf_counter()
            eng.solve(c["state"], c["target"], aux=aux_h, want_w=False, out=out)
            lat.append(time.perf_counter() - t1)
        e2e_s = time.perf_counter() - t0
        e2e_conv = int((out["status"] == 0).sum())
        e2e_value = e2e_conv * Ke / e2e_s
        h2d = B * (6 + 6 + 4) * 8
        d2h = B * (2 + 1) * 8 + B * 2 * 4
        e2e_path = ("dart_solve_host: host arrays -> pinned staging block; this batch is < 1 MB, so the kernel reads its inputs from and writes "
                    "its results to that block over the host link itself (mapped pinned memory, no copy-engine launches); sync; copy out")
    else:
        pin = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (c["state"], c["target"], aux_h)]
        host_rows = torch.empty((world * B, 4), dtype=torch.float64).pin_memory()
        eng.set_result_rows(rows)

        def e2e_step():
            for dst, src in zip((x0_d, tg_d, aux_d), pin):
                dst.copy_(src, non_blocking=True)
            eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
            if peer is not None:
                peer.handshake()
                host_rows.copy_(peer.gathered, non_blocking=True)
            else:
                dist.all_gather_into_tensor(gathered, rows)
                host_rows.copy_(gathered, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(3):
            e2e_step()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            t1 = time.perf_counter()
            e2e_step()
            lat.append(time.perf_counter() - t1)
        e2e_s = time.perf_counter() - t0
        e2 = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(e2, op=dist.ReduceOp.MAX)
        e2e_conv = int((host_rows[:, 3] == 0).sum().item())          # all ranks' rows
        e2e_value = e2e_conv * Ke / float(e2.item())
        h2d = B * (6 + 6 + 4) * 8
        d2h = world * B * 4 * 8
        e2e_path = ("pinned H2D of the rank's inputs, dart_solve, gather of the result rows (" +
                    ("peer-memory stores + hand-shake" if peer is not None else "NCCL all_gather") + "), D2H of all ranks' rows")
        eng.set_result_rows(None)
        if peer is not None:
            if int(peer.timed_out.item()) != 0:
                raise SystemExit("peer hand-shake timed out")
            peer.close()

    # ---- BASELINE config 5 scale sweep (every N; strong scaling)
    sweep = None
    if not args.no_sweep:
        try:
            sweep = scale_sweep(torch, dist, dart_b200, dev, local, world, rank, args.sweep_total, peak_tf)
        except Exception as e:      # never lose the headline line to the sub-record
            sweep = {"error": repr(e)[:300]}
            if world > 1:
                raise

    clk = clocks.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- single-instance latency (B = 1) through the host API
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
        a_, b_ = _ev(torch), _ev(torch)
        a_.record()
        for _ in range(3):
            eng.solve_device(xb, tb, aux=ab, u0_out=ub, J_out=Jb, status=sb, iters=ib)
        b_.record()
        torch.cuda.synchronize()
        sec = a_.elapsed_time(b_) * 1e-3 / 3
        itb = int(ib.sum().item())
        tfb = itb * PMPC_FLOPS_PER_ITER / sec / 1e12
        big = {"instances": Bb, "solves_per_s": float((sb == 0).sum().item()) / sec, "ms": sec * 1e3,
               "fp64_tflops": tfb, "frac": tfb / peak_tf if peak_tf else None, "launch": eng.last_launch_config()}
    except Exception as e:      # the variant is informative only; never fail the bench line on it
        big = {"error": str(e)[:200]}
    eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
    torch.cuda.synchronize()
    launch_cfg = eng.last_launch_config()

    # ---- the same batch under the monotone barrier schedule (the previous rounds' method): more iterations, each a
    # little cheaper -- shows that the roofline fraction below fell because flops were removed, not because they got slower
    mono = None
    if not os.environ.get("DART_BARRIER_MONOTONE"):
        try:
            eng.set_barrier_strategy("monotone")
            for _ in range(3):
                eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
            torch.cuda.synchronize()
            mm = []
            for _ in range(10):
                flush.zero_()
                a_, b_ = _ev(torch), _ev(torch)
                a_.record()
                eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
                b_.record()
                torch.cuda.synchronize()
                mm.append(a_.elapsed_time(b_))
            m_ms = float(np.mean(mm))
            m_it = int(it.sum().item())
            m_tf = m_it * 66.9e3 / (m_ms * 1e-3) / 1e12
            mono = {"what": "same batch, dart_set_barrier_strategy(MONOTONE)", "kernel_ms": m_ms, "mean_iters": m_it / B,
                    "solves_per_s": float((st == 0).sum().item()) / (m_ms * 1e-3), "flops_per_iteration": 66.9e3,
                    "achieved": m_tf, "frac": m_tf / peak_tf if peak_tf else None}
        except Exception as e:
            mono = {"error": repr(e)[:200]}
        finally:
            eng.set_barrier_strategy("mehrotra")
            eng.solve_device(x0_d, tg_d, aux=aux_d, u0_out=u0, J_out=J, status=st, iters=it)
            torch.cuda.synchronize()

    # ---- roofline of the solve kernel (the only kernel of the step at N = 1)
    kern_s = kern_ms * 1e-3
    flops = iters_sum * PMPC_FLOPS_PER_ITER
    achieved = flops / kern_s / 1e12
    traffic, traffic_src = _ncu_traffic()
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured in this run (dart_measure_fp64_tflops DFMA microbenchmark); MEASURED_PEAKS.json "
                               "has no FP64 entry",
                "kernel": f"nmpc_solve_kernel<PmpcAxis,{launch_cfg['lanes']},15>", "kernel_ms": kern_ms, "algorithmic_flops_per_launch": flops,
                "mean_iters": iters_sum / B, "flops_per_iteration": PMPC_FLOPS_PER_ITER,
                "barrier_strategy": "monotone" if os.environ.get("DART_BARRIER_MONOTONE") else "mehrotra predictor-corrector",
                "monotone_variant": mono,
                "note": "latency-bound: 1152 instances occupy a fraction of the SMs; see "
                                                      "throughput_variant for the filled-GPU figure"}

    cpu = None
    if not args.no_cpu_baseline:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        r, Bc, dt = cpu_oracle_rate(args.cpu_sample)
        cpu = {"value": r, "unit": "solves/s", "cores": 1, "kind": "port",
               "sample": f"oracle/ipm.py (numpy port; CasADi/IPOPT not installable) on {Bc} instances of the same batch in one "
                         f"batched call, {dt:.1f} s; reference README quotes 80-100 solves/s per IPOPT worker (PMPC/README.md:266)"}

    configs = None
    if world == 1 and not args.no_configs:
        configs = {}
        for name, fn in (("3", config3_record), ("4", config4_record)):
            try:
                configs[name] = fn(torch, dart_b200, dev, local, peak_tf)
            except Exception as e:
                configs[name] = {"error": repr(e)[:300]}
        if sweep is not None:
            configs["5"] = {"see": "scale_sweep", "solves_per_s": sweep.get("solves_per_s"), "converged": sweep.get("converged"),
                            "roofline": sweep.get("roofline")}

    line = {"metric": "NMPC solves/sec", "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_s_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PMPC batched (BASELINE config 2): 18 objects x {STATES_PER_OBJECT} states = {B} "
                                   f"instances/step/GPU, cold start, tol 1e-8", "N": 15, "instances_per_gpu": B,
                       "l2": "256 MiB flush write between timed steps", "launch": launch_cfg,
                       "parallelism": f"instance sharding x{world}" + (", one gather of the result rows per step INSIDE the timed events" if world > 1 else ""),
                       "gather": gather_kind},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "p50_batch_latency_ms": float(np.median(lat) * 1e3), "path": e2e_path},
            "p50_solve_latency_ms": {"B=1 host API": float(np.median(one) * 1e3),
                                     "per-batch/B": float(np.median(lat) * 1e3 / B)},
            "gpu_launches": int(launches), "converged": conv, "gather_ms": gather_ms, "nccl_gather_ms": nccl_gather_ms, "clocks": clk, "roofline": roofline,
            "cpu_baseline": cpu, "throughput_variant": big, "scale_sweep": sweep, "configs": configs}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
