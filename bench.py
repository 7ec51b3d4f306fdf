#!/usr/bin/env python
"""Benchmark of the render + fitness hot path (BASELINE.json metric: candidate evals/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one fused render+fitness evaluation of one population on each GPU (plus, for N > 1,
the NCCL all-gather of the fitness vector).  Default workload is BASELINE config 3:
256x256 px, 1,000 splats, population 1,024 per GPU, mask-weighted fitness.  One JSON line on
stdout (rank 0); everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "genetic-gaussian-splats_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "candidate_evals_per_sec"
UNIT = "candidates/s"
FLOP_PER_PAIR = 23          # SURVEY.md section 8d, counted from render.py:189-196
FLOP_PER_PIXEL = 12         # fitness.py:16-31
NOMINAL_FP32_TFLOPS = 74.4  # 148 SM x 128 lanes x 2 flop x 1.965 GHz (clocks.max.sm)

WORKLOADS = {
    # name: H, W, N splats, population per GPU, masked fitness
    "c1": dict(H=128, W=128, N=100, P=32, mask=True,
               desc="config 1: 128x128, 100 splats, population 32"),
    "c2": dict(H=256, W=256, N=500, P=8, mask=True,
               desc="config 2: 256x256, 500 splats, 8 batched SA neighbours"),
    "c3": dict(H=256, W=256, N=1000, P=1024, mask=True,
               desc="config 3: GA 256x256, 1,000 splats, population 1,024 per GPU, "
                    "mask-weighted fitness"),
    "c4": dict(H=512, W=512, N=4000, P=1024, mask=True,
               desc="config 4 shard: 512x512, 4,000 splats, 1,024 candidates per GPU "
                    "(8,192 over 8 GPUs)"),
}
POOL = 4  # distinct populations rotated through the timed loop


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Libraries (NCCL prints its version banner) may write to fd 1; the contract is ONE JSON line
# on stdout, so everything but that line is sent to stderr at the file-descriptor level.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def ncu_traffic(workload: str, population: int):
    """dram bytes read + written per raster launch, from the committed ncu capture of this
    workload (profiles/rNN_traffic.json); None when no capture matches."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for name in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if name.endswith("_traffic.json"):
            try:
                rec = json.load(open(os.path.join(pdir, name)))
            except Exception:
                continue
            if rec.get("workload") == workload and rec.get("population") == population:
                best = rec["dram_bytes_read"] + rec["dram_bytes_write"]
    return best


def reference_gpu_recorded(workload: str):
    """The reference's real Triton path on a B200, as RECORDED by tools/reference_gpu_compare.py
    (profiles/rNN_reference_gpu_compare.json): it cannot run inside bench.py (the reference is
    not on the GPU box).  Context for the reader, not a live measurement."""
    tag = {"c1": "config 1", "c2": "config 2", "c3": "config 3"}.get(workload)
    pdir = os.path.join(ROOT, "profiles")
    found = None
    for name in sorted(os.listdir(pdir)) if (tag and os.path.isdir(pdir)) else []:
        if name.endswith("_reference_gpu_compare.json"):
            try:
                rec = json.load(open(os.path.join(pdir, name)))
            except Exception:
                continue
            for row in rec.get("rows", []):
                if row.get("shape", "").startswith(tag):
                    found = {"value": row["reference_candidates_per_s"], "unit": UNIT,
                             "ms_per_call": row["reference_ms_per_call"],
                             "candidates_per_call": row["candidates"], "source": "profiles/" + name,
                             "what": "unmodified reference (Triton) through fitness_population on "
                                     "one B200, recorded run, wall clock"}
    return found


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


# --------------------------------------------------------------------------- clocks

class ClockSampler:
    """SM clock, power and throttle reasons sampled WHILE the timed region runs: NVML polled
    from a thread every few milliseconds (nvidia-smi once as a fallback)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, uuid: str, index: int):
        self.uuid, self.index = uuid, index
        self.samples, self.stop_flag, self.thread, self.h = [], threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
        except Exception as e:
            log(f"[bench] NVML unavailable ({e}); falling back to one nvidia-smi query")

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self) -> dict:
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
        if self.samples:
            sm = [s[0] for s in self.samples]
            mask = 0
            for s_ in self.samples:
                mask |= int(s_[2])
            try:
                mx = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            except Exception:
                mx = max(sm)
            return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(mx),
                    "power_w_max": float(max(s[1] for s in self.samples)),
                    "reasons": sorted(n for bit, n in self.REASONS.items() if mask & bit),
                    "samples": len(sm), "how": "NVML polled every ~4 ms during the timed region"}
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index),
                                  "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                 timeout=20).stdout.strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]),
                    "power_w_max": float(out[2]), "reasons": [], "samples": 1,
                    "how": "one nvidia-smi query after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


# ------------------------------------------------------------------ CPU (reference) arm

def cpu_sample(wl, seconds: float, rank_seed: int = 0):
    """Times the CPU oracle (the restatement of the reference's arithmetic; the reference has
    no CPU path of its own) on a bounded sample of the workload.  Returns (cands/s, sample)."""
    from ggs_b200 import synth
    from oracle import oracle
    H, W, N = wl["H"], wl["W"], wl["N"]
    cores = oracle.threads()
    t = synth.synthetic_target_np(H, W, 0)
    m = synth.importance_mask_np(t) if wl["mask"] else None
    probe = synth.new_population_np(cores, N, H, W, seed=42 + rank_seed)
    t0 = time.perf_counter()
    oracle.fitness(probe, t, H, W, 3.0, weight_mask=m)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(max(cores, min(wl["P"], round(seconds / dt) * cores)))
    g = synth.new_population_np(n, N, H, W, seed=43 + rank_seed)
    t0 = time.perf_counter()
    oracle.fitness(g, t, H, W, 3.0, weight_mask=m)
    dt = time.perf_counter() - t0
    return n / dt, n, cores, dt


def run_reference(args, wl):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from oracle import oracle
    oracle.build()
    cores = oracle.threads()
    per_step = max(0.5, min(20.0, 120.0 / max(1, args.steps + args.warmup)))  # ~2 min in all
    for _ in range(args.warmup):
        cpu_sample(wl, per_step / 4)
    vals, n_last, t_total = [], 0, 0.0
    for s in range(args.steps):
        v, n, cores, dt = cpu_sample(wl, per_step, rank_seed=s)
        vals.append(v)
        n_last = n
        t_total += dt
    value = float(np.mean(vals))
    sample = (f"{n_last} candidates of {wl['H']}x{wl['W']}/{wl['N']} splats per step, "
              f"CPU oracle (oracle/ggs_oracle.c, OpenMP over candidates)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "H": wl["H"], "W": wl["W"], "splats": wl["N"],
                   "population_per_gpu": wl["P"], "fitness": "mask" if wl["mask"] else "plain",
                   "note": "the reference has no CPU path (render.py:217 asserts CUDA); this arm "
                           "times the CPU restatement of its arithmetic on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------ our arm

def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    import ggs_b200
    from ggs_b200 import synth

    rank, local_rank, world = dist_env()
    if args.gpus > 1 and world != args.gpus:
        log(f"[bench] --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
        return 2
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    ggs_b200.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    H, W, N, P = wl["H"], wl["W"], wl["N"], wl["P"]
    K, Wm = args.steps, args.warmup
    t_np = synth.synthetic_target_np(H, W, 0)
    m_np = synth.importance_mask_np(t_np) if wl["mask"] else None
    target = torch.from_numpy(t_np).to(dev)
    mask = None if m_np is None else torch.from_numpy(m_np).to(dev)

    # Each rank owns its shard of the population (weak scaling: P candidates per GPU).  POOL
    # distinct populations are rotated so consecutive steps never read the same genomes.
    host_pool = [torch.from_numpy(synth.new_population_np(P, N, H, W, seed=42 + 1000 * rank + r))
                 .pin_memory() for r in range(POOL)]
    pool = [h.to(dev) for h in host_pool]
    pool_bytes = sum(h.numel() * 4 for h in host_pool)

    pairs = []
    for g in pool:
        d = ggs_b200.decode(g, H, W, 3.0, layout=ggs_b200.LAYOUT_AXES_ANGLE)
        area = (d["x1"] - d["x0"] + 1).clamp_min(0).long() * (d["y1"] - d["y0"] + 1).clamp_min(0).long()
        pairs.append(int(area.sum().item()))
    flop_per_launch = [FLOP_PER_PAIR * p + FLOP_PER_PIXEL * P * H * W for p in pairs]

    peaks = ggs_b200.probe_peaks()
    log(f"[bench rank {rank}] probe: {peaks}")

    all_fit = torch.empty((world * P,), dtype=torch.float32, device=dev) if world > 1 else None

    def step(i):
        fit = ggs_b200.fitness(pool[i % POOL], target, H, W, 3.0, weight_mask=mask)
        if world > 1:
            dist.all_gather_into_tensor(all_fit, fit)
            return all_fit
        return fit

    for i in range(max(Wm, 0)):
        step(i)
    torch.cuda.synchronize(dev)

    props = torch.cuda.get_device_properties(dev)
    sampler = ClockSampler("GPU-" + str(props.uuid) if hasattr(props, "uuid") else "", local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ggs_b200.timing_enable(True)
    sampler.start()
    e0.record()
    last = None
    for i in range(K):
        last = step(Wm + i)
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    kt = ggs_b200.timing_read()
    ggs_b200.timing_enable(False)
    clocks = sampler.stop()
    assert torch.isfinite(last).all()

    # ---- work actually done (instrumented kernel, outside the timed region)
    work = ggs_b200.count_evaluated_pairs(
        lambda: ggs_b200.fitness(pool[0], target, H, W, 3.0, weight_mask=mask), device=dev)

    # ---- end to end: host buffers through the C ABI, H2D + D2H inside the timed region
    he = ggs_b200.HostEvaluator(t_np, m_np, device=local_rank)
    out = torch.empty((P,), dtype=torch.float32).pin_memory()
    for i in range(min(max(Wm, 1), 3)):
        he.fitness(host_pool[i % POOL], out=out)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        he.fitness(host_pool[(Wm + i) % POOL], out=out)
    e2e_s = time.perf_counter() - t0
    he.close()
    # the host path must agree bit for bit with the device path on the same genomes
    chk = ggs_b200.fitness(pool[(Wm + K - 1) % POOL], target, H, W, 3.0, weight_mask=mask)
    assert torch.equal(chk.cpu(), out), "host path disagrees with device path"

    t_max = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all = float(t_max[0]), float(t_max[1])

    if rank == 0:
        value = world * P * K / (ms_all * 1e-3)
        e2e_value = world * P * K / (e2e_ms_all * 1e-3)
        flops_timed = sum(flop_per_launch[(Wm + i) % POOL] for i in range(K))
        raster_s = kt["raster_ms"] * 1e-3
        achieved = flops_timed / raster_s / 1e12 if raster_s > 0 else None
        roofline = {
            "bound": "fp32", "kernel": "ggs::raster_kernel", "achieved": achieved,
            "peak": NOMINAL_FP32_TFLOPS, "unit": "TFLOP/s",
            "frac": None if achieved is None else achieved / NOMINAL_FP32_TFLOPS,
            "traffic": ncu_traffic(args.workload, P),
            "peak_source": "nominal 148 SM x 128 lanes x 2 x clocks.max.sm 1965 MHz; "
                           "MEASURED_PEAKS.json has no fp32 entry (HBM and bf16 tensor only)",
            "measured_ffma_tflops": peaks["ffma_tflops"],
            "measured_ffma2_tflops": peaks["ffma2_tflops"],
            "measured_mufu_ex2_gops": peaks["mufu_ex2_gops"],
            "frac_of_measured_ffma": None if achieved is None else achieved / peaks["ffma_tflops"],
            "algorithmic_flop_per_launch": float(np.mean(flop_per_launch)),
            "pairs_per_candidate": float(np.mean(pairs)) / P,
            "evaluated_pairs_per_candidate": work["pairs"] / P,
            "evaluated_over_algorithmic": work["pairs"] / max(1, pairs[0]),
            "evaluated_on_exact_path": work["exact_pairs"] / max(1, work["pairs"]),
            "note": "evaluated pairs = lanes x rows the kernel really blended (instrumented run): "
                    "above the algorithmic count by the columns of a 32-wide tile outside a "
                    "splat's AABB, below it where bands stop once opaque (transmittance < 2^-22)",
            "raster_ms_per_launch": kt["raster_ms"] / max(1, kt["evaluations"]),
            "decode_ms_per_launch": kt["decode_ms"] / max(1, kt["evaluations"]),
            "raster_share_of_step": kt["raster_ms"] / ms if ms > 0 else None,
        }
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                v, n, cores, dt = cpu_sample(wl, args.cpu_seconds)
                cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": f"{n} candidates of the same workload in {dt:.1f} s, CPU oracle "
                                 f"(oracle/ggs_oracle.c, OpenMP over candidates, scalar expf)"}
            except Exception as e:
                log(f"[bench] cpu_baseline failed: {e}")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_all / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "H": H, "W": W, "splats": N,
                       "population_per_gpu": P, "population_total": world * P,
                       "fitness": "mask" if wl["mask"] else "plain", "k_sigma": 3.0,
                       "parallelism": f"population sharded over {world} GPU(s), NCCL all-gather of "
                                      f"the fitness vector" if world > 1 else "1 GPU",
                       "l2": f"inputs rotate over {POOL} populations ({pool_bytes / 1e6:.0f} MB "
                             f"> 126 MB L2)" if pool_bytes > 126e6 else
                             f"inputs rotate over {POOL} populations ({pool_bytes / 1e6:.0f} MB); "
                             f"compute-bound, {N * 36} B of genome per candidate"},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": P * N * 9 * 4, "d2h_bytes_per_step": P * 4,
                    "ms_per_step": e2e_ms_all / K,
                    "api": "ggs_ctx_fitness_host (C ABI, pinned host genomes in, host fitness out)"},
            "gpu_launches": 2 * K,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "reference_gpu_recorded": reference_gpu_recorded(args.workload) if world == 1 else None,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c3")
    ap.add_argument("--population", type=int, default=None, help="override candidates per GPU")
    ap.add_argument("--side", type=int, default=None, help="override H = W (roofline sweep)")
    ap.add_argument("--splats", type=int, default=None, help="override splats per candidate")
    ap.add_argument("--pool", type=int, default=None, help="distinct populations rotated (default 4)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.population:
        wl["P"] = args.population
    if args.side or args.splats:
        wl["H"] = wl["W"] = args.side or wl["H"]
        wl["N"] = args.splats or wl["N"]
        wl["desc"] = (f"sweep point: {wl['H']}x{wl['W']}, {wl['N']} splats, population {wl['P']} "
                      f"per GPU, mask-weighted fitness")
    if args.pool:
        global POOL
        POOL = args.pool
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: fewer than 3 warm-up steps requested")
    return run_reference(args, wl) if args.impl == "reference" else run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
