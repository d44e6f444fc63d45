#!/usr/bin/env python
"""bench.py - BASELINE.json metric: QP solves/s, batch 4096 Lite3 trot problems (N=10) per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (condense + factor + ADMM, cold start) over one batch
of 4096 synthetic problems per GPU.  One process per GPU
(torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); problems shard by batch index with no
collective on the solve path ("weak" scaling: 4096 problems per GPU).  Rank 0 prints one
JSON line.  `--impl reference` times the reference's CPU algorithm (OSQP restatement in C,
oracle/) on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BATCH = 4096
HORIZON = 10
KERNEL_NAMES = {10: "cmpc::solve_kernel<10,1,8,1,false>", 20: "cmpc::solve_kernel<20,2,2,1,false>",
                30: "cmpc::solve_kernel<30,6,1,3,false>", 40: "cmpc::solve_cluster_kernel<10,4,4,1>",
                60: "cmpc::solve_cluster_kernel<10,6,6,1>"}
WORKLOAD = "config2: batch 4096 Lite3 trot MPC QPs per GPU, N=10, randomized CoM states/velocity refs, cold start"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--horizon", type=int, default=HORIZON)
    ap.add_argument("--gaits", default="trot")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thr.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def algorithmic_work(N, stance_counts, iters):
    """SURVEY.md section 8(d) work model per solve (averaged over the batch):
    FLOPs = R_f (n^3/3 + 2 n^2) + 12 n^2 + K (2.08 n^2 + 33 n), smem bytes = K 4 n^2,
    HBM bytes = 4 (62 N + 44) + N."""
    n = 3.0 * stance_counts.astype(np.float64)
    K = iters.astype(np.float64)
    flops = (n ** 3 / 3 + 2 * n ** 2) + 12 * n ** 2 + K * (2.08 * n ** 2 + 33 * n)
    smem = K * 4 * n ** 2
    hbm = 4.0 * (62 * N + 44) + N
    return float(flops.mean()), float(smem.mean()), float(hbm)


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's CPU algorithm (fp64 OSQP restatement, oracle/) on the
    host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline
    import mpc_b200 as pkg
    sample = args.cpu_sample or 256
    pb = pkg.problems.synthetic_batch(sample, N=args.horizon, gaits=tuple(args.gaits.split(",")), seed=0)
    for _ in range(max(args.warmup, 1)):
        cpu_baseline.solve_batch(pb.slice(0, min(32, sample)))
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        info = cpu_baseline.solve_batch(pb)
        times.append(time.perf_counter() - t0)
    tot = sum(times)
    val = sample * args.steps / tot
    line = {
        "impl": "reference", "metric": "QP solves/sec", "value": val, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample} problems per step"},
        "cpu_baseline": {"value": val, "unit": "solves/s", "cores": info["threads"],
                         "kind": info["kind"], "sample": f"{sample} of the {args.batch} problems, "
                         f"all {info['threads']} host threads, mean {info['mean_iters']:.1f} OSQP iterations"},
        "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mpc_b200 as pkg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N = args.batch, args.horizon
    gaits = tuple(args.gaits.split(","))
    # every rank draws its own shard of the global batch (batch-index sharding, no exchange)
    pb = pkg.problems.synthetic_batch(B, N=N, gaits=gaits, seed=rank)
    host = pb.f32()
    pinned = [torch.from_numpy(a).pin_memory() for a in host]
    dargs = [t.to(dev) for t in pinned]
    mpc = pkg.BatchedMPC(N=N, max_batch=B, device=local, warm_mode=0)
    out = mpc.alloc_outputs(B, want_X=True, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stream = torch.cuda.current_stream(dev)
    for _ in range(max(args.warmup, 3)):
        mpc.solve(*dargs, out=out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = mpc.launch_count
    # --- kernel-resident timing: inputs already in HBM, per-step CUDA events, L2 flushed ----
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                      # evict inputs / warm state from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mpc.solve(*dargs, out=out)
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    launches = mpc.launch_count - l0
    dev_ms = sum(step_ms)
    iters = out[2].cpu().numpy()
    status = out[5].cpu().numpy()

    # --- the solve kernel alone (roofline denominator): CUDA events recorded by the library on the
    # launching stream immediately around that one kernel, same inputs, L2 flushed between steps
    mpc_t = pkg.BatchedMPC(N=N, max_batch=B, device=local, warm_mode=0, time_kernel=1)
    kern_ms = []
    for i in range(args.steps + 3):
        flush.fill_(1)
        mpc_t.solve(*dargs, out=out)
        if i >= 3:
            kern_ms.append(mpc_t.last_kernel_ms)
    barrier()
    mpc_t.close()

    # --- warm start: x of the same problems one tick earlier, y = 0 (the
    # reference's semantics, src/mpc.py:270-271).  The previous-tick solve and the restore of the
    # warm state are outside the timed events; only the warm-started solve is timed.
    pb_prev = pkg.problems.synthetic_batch(B, N=N, gaits=gaits, seed=rank, tick_shift=-1)
    prev_args = [torch.from_numpy(a).to(dev) for a in pb_prev.f32()]
    mpc_w = pkg.BatchedMPC(N=N, max_batch=B, device=local, warm_mode=1)
    out_prev = mpc_w.alloc_outputs(B, want_X=False, device=dev)
    out_w = mpc_w.alloc_outputs(B, want_X=True, device=dev)
    mpc_w.solve(*prev_args, want_X=False, out=out_prev)
    x_prev = out_prev[0].clone()
    warm_ms = []
    for i in range(args.steps + 3):
        mpc_w.set_warm(x_prev)
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mpc_w.solve(*dargs, out=out_w)
        e1.record(stream)
        e1.synchronize()
        if i >= 3:
            warm_ms.append(e0.elapsed_time(e1))
    warm_iters = out_w[2].cpu().numpy()
    warm_status = out_w[5].cpu().numpy()
    barrier()
    mpc_w.close()

    # --- end to end: host (pinned) buffers through cmpc_solve_host, copies inside the timing ----
    hin = [t.numpy() for t in pinned]
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    hout = (pin((B, N, 12), torch.float32), None, pin((B,), torch.int32),
            pin((B,), torch.float32), pin((B,), torch.float32), pin((B,), torch.int32))
    for _ in range(3):
        mpc.solve_host(*hin, want_X=False, out=hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mpc.solve_host(*hin, want_X=False, out=hout)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # --- latency: B=1 end-to-end p50, batch wall p50 ----------------------------------------
    lat = {}
    if rank == 0:
        one = [t[:1].contiguous() for t in dargs]
        out1 = mpc.alloc_outputs(1, want_X=True, device=dev)
        ts = []
        for i in range(220):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            mpc.solve(*one, out=out1)
            e1.record(stream)
            e1.synchronize()
            if i >= 20:
                ts.append(e0.elapsed_time(e1))
        lat = {"b1_solve_p50_us": 1e3 * statistics.median(ts),
               "batch_p50_ms": statistics.median(step_ms),
               "batch_amortised_p50_us_per_solve": 1e3 * statistics.median(step_ms) / B}

    # --- reduce over ranks: max time ------------------------------------------------------------
    tt = torch.tensor([dev_ms, e2e_s, sum(warm_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max, warm_ms_max = float(tt[0]), float(tt[1]), float(tt[2])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total = B * world * args.steps
    value = total / (dev_ms_max * 1e-3)
    e2e_val = total / e2e_s_max
    hbm_peak, peak_src = measured_peaks()
    flops, smem_b, hbm_b = algorithmic_work(N, pb.stance.sum((1, 2)), iters)
    kern_s = statistics.mean(kern_ms) * 1e-3         # this rank's mean solve-kernel launch duration
    ach_gbs = hbm_b * B / kern_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and (B, N, gaits) == (BATCH, HORIZON, ("trot",)):   # captured on this workload only
        traffic = json.load(open(tp)).get("solve_kernel_dram_bytes_per_launch")
    try:
        fp32_peak = pkg._capi.fp32_peak(local) * 1e12
        fp32_src = "measured in this run (cmpc_fp32_peak: 8-chain FMA kernel, best of 4)"
    except Exception:
        fp32_peak = 148 * 128 * 2 * 1.965e9
        fp32_src = "148 SM x 128 lanes x 2 x 1.965 GHz (nominal)"
    smem_peak = 148 * 128 * 1.965e9
    roof = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": ach_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
            "kernel": KERNEL_NAMES.get(N, f"cmpc::solve_kernel<{N},...>"), "algorithmic_bytes_per_solve": hbm_b,
            "kernel_ms": statistics.mean(kern_ms), "kernel_share_of_step": statistics.mean(kern_ms) / (dev_ms / args.steps),
            "timing": "cudaEvent pair recorded by libcmpc on the launching stream around the solve kernel only "
                      "(cfg.time_kernel), mean over the timed steps of a second pass with the same inputs",
            "note": "HBM is NOT the binding roof of this kernel (SURVEY.md 8d): on-chip fractions follow",
            "fp32": {"algorithmic_flop_per_solve": flops, "achieved_tflops": flops * B / kern_s / 1e12,
                     "peak_tflops": fp32_peak / 1e12, "frac": flops * B / kern_s / fp32_peak,
                     "peak_source": fp32_src},
            "smem": {"algorithmic_bytes_per_solve": smem_b, "achieved_tbs": smem_b * B / kern_s / 1e12,
                     "peak_tbs": smem_peak / 1e12, "frac": smem_b * B / kern_s / smem_peak}}
    in_bytes = sum(a.nbytes for a in host)
    out_bytes = sum(a.nbytes for a in hout if a is not None)
    line = {
        "metric": "QP solves/sec", "value": value, "unit": "solves/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD if (B, N, gaits) == (BATCH, HORIZON, ("trot",)) else
                   f"batch {B} x N={N} gaits={','.join(gaits)} per GPU, cold start",
                   "batch_per_gpu": B, "horizon": N, "gaits": list(gaits), "sharding": f"batch-index x{world}",
                   "l2": "flushed between steps (256 MiB fill), per-step CUDA-event timing",
                   "solver": {"rho": float(mpc.cfg.rho), "sigma": float(mpc.cfg.sigma),
                              "alpha": float(mpc.cfg.alpha), "eps_abs": float(mpc.cfg.eps_abs),
                              "eps_rel": float(mpc.cfg.eps_rel), "check_every": int(mpc.cfg.check_every)},
                   "mean_iters": float(iters.mean()), "max_iters": int(iters.max()),
                   "solved_frac": float((status == 1).mean())},
        "roofline": roof,
        "e2e": {"value": e2e_val, "unit": "solves/s", "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": out_bytes,
                "api": "cmpc_solve_host, page-locked host buffers " +
                       ("read / written in place by the solve kernel over PCIe (host_zero_copy)"
                        if int(mpc.cfg.host_zero_copy) else "staged with chunked cudaMemcpyAsync")},
        "warm_start": {"value": total / (warm_ms_max * 1e-3), "unit": "solves/s",
                       "ms_per_step": warm_ms_max / args.steps, "mean_iters": float(warm_iters.mean()),
                       "max_iters": int(warm_iters.max()), "solved_frac": float((warm_status == 1).mean()),
                       "what": "same batch warm-started with the forces of the same problems one tick earlier "
                               "(x unshifted, y = 0: the reference's set_initial semantics), L2 flushed, CUDA events"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "latency": lat,
        "wall_s_timed_region": t_wall,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            from oracle import cpu_baseline
            sample = args.cpu_sample or 192
            info = cpu_baseline.solve_batch(pb.slice(0, sample))
            line["cpu_baseline"] = {
                "value": sample / info["seconds"], "unit": "solves/s", "cores": info["threads"],
                "kind": info["kind"],
                "sample": f"first {sample} of the {B} problems, {info['threads']} host threads, "
                          f"mean {info['mean_iters']:.1f} OSQP iterations"}
        except Exception as exc:      # the oracle is a checker; a missing build must not kill the bench
            line["cpu_baseline"] = {"value": None, "unit": "solves/s", "cores": 0, "kind": "port",
                                    "sample": f"unavailable: {exc}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
