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


def workload_config(wl, gpus: int) -> dict:
    """The `config` object of the JSON line: a function of the workload and the GPU count only,
    so both arms (`--impl ours` / `--impl reference`) print the same one."""
    P, N = wl["P"], wl["N"]
    pool_bytes = POOL * P * N * 9 * 4
    return {"workload": wl["desc"], "H": wl["H"], "W": wl["W"], "splats": N,
            "population_per_gpu": P, "population_total": gpus * P,
            "fitness": "mask" if wl["mask"] else "plain", "k_sigma": 3.0,
            "parallelism": (f"population sharded over {gpus} GPU(s), fitness vector gathered on "
                            f"every rank") if gpus > 1 else "1 GPU",
            "l2": (f"inputs rotate over {POOL} populations ({pool_bytes / 1e6:.0f} MB > 126 MB L2)"
                   if pool_bytes > 126e6 else
                   f"inputs rotate over {POOL} populations ({pool_bytes / 1e6:.0f} MB); "
                   f"compute-bound, {N * 36} B of genome per candidate")}


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_gpu_available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "modules"))


def reference_gpu_live(args, wl, device_index: int, budget_s: float = 25.0):
    """The UNMODIFIED reference (baseline/_ref, git-ignored copy of /root/reference) through its
    stock path -- prewarm_renderer, then fitness_population -> fitness_many ->
    render_splats_rgb_triton (its Triton kernel, JIT-compiled for this GPU) -- on the same
    genomes, target and workload, timed live in a child process (both trees call their package
    `modules`).  Returns (dict for the JSON line, path of an .npz with its mask and fitness)."""
    if not reference_gpu_available():
        return {"unavailable": "baseline/_ref missing (run baseline/make_ref_copy.sh in the build "
                               "container; /root/reference does not exist on the GPU box)"}, None
    import tempfile
    out_npz = os.path.join(tempfile.mkdtemp(prefix="ggs_ref_"), "ref.npz")
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-gpu-child",
           "--workload", args.workload, "--population", str(wl["P"]), "--side", str(wl["H"]),
           "--splats", str(wl["N"]), "--ref-out", out_npz, "--ref-budget", str(budget_s)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "")
               .split(",")[device_index] if os.environ.get("CUDA_VISIBLE_DEVICES") else str(device_index))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            return {"unavailable": "reference child failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:300]}, None
        return json.loads(line[-1]), out_npz
    except Exception as e:  # timeout, ...
        return {"unavailable": f"reference child: {type(e).__name__}: {e}"[:300]}, None


def run_reference_gpu_child(args, wl):
    """Child process of reference_gpu_live: nothing of this repository's product is imported."""
    import importlib.util
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, PKG)]
    sys.path.insert(0, REF_DIR)
    import torch
    import modules.fitness as rfit
    import modules.mask as rmask
    import modules.utils as rutils
    assert os.path.abspath(rfit.__file__).startswith(REF_DIR), rfit.__file__
    spec = importlib.util.spec_from_file_location("ggs_synth", os.path.join(PKG, "ggs_b200", "synth.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)           # numpy only: seeded inputs, not product code
    H, W, N, P = wl["H"], wl["W"], wl["N"], wl["P"]
    t_np = synth.synthetic_target_np(H, W, 0)
    target = torch.from_numpy(t_np).cuda()
    mask = None
    if wl["mask"]:                           # algorithm.py:42-49
        mask = rmask.compute_importance_mask(target, H, W, edge_scales=(1, 2, 4), w_edge=0.7,
                                             w_var=0.3, gamma=0.7, floor=0.15, smooth=3,
                                             strength=0.7).to("cuda")
    pop = list(torch.from_numpy(synth.new_population_np(P, N, H, W, seed=42)).cuda().unbind(0))
    rutils.prewarm_renderer(H, W, 3.0, "cuda")      # algorithm.py:52

    def call():
        return rfit.fitness_population(pop, target, H, W, 3.0, "cuda", tile=32, chunk=None,
                                       weight_mask=mask, boost_only=False)
    fit = call()                              # warm-up: Triton JIT for this shape
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit = call()
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    reps = int(max(1, min(20, args.ref_budget / max(first, 1e-4))))
    times = [first]
    for _ in range(reps - 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fit = call()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    if args.ref_out:
        np.savez(args.ref_out, fitness=np.asarray(fit, dtype=np.float64),
                 mask=(mask.cpu().numpy() if mask is not None else np.zeros(0, np.float32)))
    import triton
    best, mean = min(times), sum(times) / len(times)
    emit({"value": P / mean, "unit": UNIT, "candidates": P, "ms_per_call": mean * 1e3,
          "best_ms_per_call": best * 1e3, "calls_timed": len(times),
          "what": "unmodified reference (baseline/_ref = /root/reference/modules) through "
                  "prewarm_renderer + fitness_population(list) -> List[float] "
                  "(fitness.py:35-48 -> render_splats_rgb_triton, Triton JIT), same GPU, same "
                  "genomes / target, wall clock around synchronised calls after one warm-up call",
          "triton": triton.__version__, "torch": torch.__version__,
          "gpu": torch.cuda.get_device_name(0)})
    return 0


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

def host_cores() -> int:
    """Host threads this process may use (its affinity mask, else the CPU count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample(wl, seconds: float, rank_seed: int = 0):
    """Times the CPU oracle (the restatement of the reference's arithmetic; the reference has
    no CPU path of its own) on a bounded sample of the workload.  Returns (cands/s, sample)."""
    from ggs_b200 import synth
    from oracle import oracle
    from oracle import torch_ref
    H, W, N = wl["H"], wl["W"], wl["N"]
    cores = host_cores()
    oracle.set_threads(cores)    # explicit: torchrun exports OMP_NUM_THREADS=1
    t = synth.synthetic_target_np(H, W, 0)
    m = torch_ref.importance_mask_np(t) if wl["mask"] else None
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
    """The reference arm.  The reference has no CPU implementation of this path (render.py:217
    asserts CUDA), so `value` is the CPU restatement of its arithmetic (oracle/ggs_oracle.c) on
    all host cores, as the task's tier rules ask; next to it, `reference_gpu` is the reference's
    REAL path -- its Triton kernel on this GPU -- timed live when baseline/_ref travelled here."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    from oracle import oracle
    oracle.build()
    cores = host_cores()
    oracle.set_threads(cores)
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
              f"CPU oracle (oracle/ggs_oracle.c, OpenMP over candidates, {cores} threads)")
    ref_gpu = None
    try:
        import torch
        if torch.cuda.is_available():
            ref_gpu, _ = reference_gpu_live(args, wl, local_rank)
    except Exception as e:
        ref_gpu = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, args.gpus),
        "note": "the reference has no CPU path (render.py:217 asserts CUDA): `value` times the CPU "
                "restatement of its arithmetic on the host cores; `reference_gpu` is the "
                "reference's own Triton path on this GPU",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_gpu": ref_gpu,
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------ our arm

def config4_leg(rank, world, dev, peers, steps=3):
    """BASELINE config 4: GA at 512x512, 4,000 splats, population 8,192 sharded over the N GPUs.
    (a) the evaluation of the whole population (each rank its 8,192 / N slice, fitness gathered
    on every rank); (b) a WHOLE generation on the device engine -- breed, evaluate the children,
    gather, elitism + ranking -- with no host sync.  Max over ranks, CUDA events."""
    import torch
    import torch.distributed as dist
    import ggs_b200
    from ggs_b200 import synth
    from ggs_b200.distributed import shard_bounds
    from ggs_b200.engine import GaEngine
    from modules.population import new_population
    from modules.utils import build_mut_sigma, scale_log_bounds
    import modules.config as C
    H = W = 512
    N, P, n_elite = 4000, 8192, 8
    t_np = synth.synthetic_target_np(H, W, 0)
    target = torch.from_numpy(t_np).to(dev)
    mask = ggs_b200.importance_mask(target, H, W, edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3,
                                    gamma=0.7, floor=0.15, smooth=3, strength=0.7)
    torch.manual_seed(42)                       # the same population on every rank
    pop = new_population(P, N, H, W, 3.0, 0.1, device=dev)
    lo, hi = shard_bounds(P, world, rank)

    def timed(fn, n):
        fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # the yardstick: 1,024 candidates of this shape on ONE GPU, no exchange
    base_ms = timed(lambda: ggs_b200.fitness(pop[lo:lo + 1024], target, H, W, 3.0, weight_mask=mask), steps)
    base_rate = 1024 / (base_ms * 1e-3)

    if world > 1 and peers is not None:
        def evaluate():
            return peers.fitness_allgather(pop[lo:hi], target, H, W, offset=lo, total=P, weight_mask=mask)
    elif world > 1:
        full = torch.empty((P,), dtype=torch.float32, device=dev)
        split = ggs_b200.choose_split(P, N, H, W)

        def evaluate():
            dist.all_gather_into_tensor(full, ggs_b200.fitness(pop[lo:hi], target, H, W, 3.0,
                                                               weight_mask=mask, split=split))
            return full
    else:
        def evaluate():
            return ggs_b200.fitness(pop, target, H, W, 3.0, weight_mask=mask)
    eval_ms = timed(evaluate, steps)
    eval_rate = P / (eval_ms * 1e-3)

    out = {"workload": "config 4: GA 512x512, 4,000 splats, population 8,192 (fixed) over N GPUs",
           "population_total": P, "population_per_gpu": hi - lo, "n_gpus": world, "scaling": "strong",
           "evaluation": {"candidates_per_s": eval_rate, "ms_per_population": eval_ms},
           "one_gpu_1024_candidates": {"candidates_per_s": base_rate, "ms": base_ms},
           "efficiency_vs_n_times_one_gpu_1024_rate": eval_rate / (world * base_rate)}

    # (b) whole generations on the engine (P - n_elite children bred, evaluated, ranked per generation)
    gens = steps
    if world == 1 or peers is not None:
        eng = GaEngine(target, mask, H, W, P, N, n_elite, 2 * gens + 2)
        if world > 1:
            eng.set_peers(peers)
        eng.start(pop, 7)
        lo_s, hi_s = scale_log_bounds(H, W, C.MIN_SCALE_SPLATS, C.MAX_SCALE_SPLATS)
        rows = [build_mut_sigma(1, 100, C.SCHEDULE, C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN)] * gens
        gen_ms = timed(lambda: eng.run(rows, C.TOUR_K, C.CXPB, C.MUTPB, lo_s, hi_s), 1) / gens
        st = eng.state(want_best=False)
        eng.close()
        out["generation"] = {"ms_per_generation": gen_ms, "generations_per_s": 1e3 / gen_ms,
                             "candidates_per_s": (P - n_elite) / (gen_ms * 1e-3),
                             "what": "breed 8,184 children (replicated on every rank) + evaluate this "
                                     "rank's slice + fitness exchange + elitism / ranking, enqueued on "
                                     "the device engine without host syncs",
                             "best_fitness_after": st["best_fitness"]}
    else:
        out["generation"] = {"unavailable": "the sharded engine needs the peer-to-peer exchange"}
    del pop
    torch.cuda.empty_cache()
    return out


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
    numa = None
    if world > 1 and args.numa_bind:
        # one process per GPU: run, and first-touch the pinned genome buffers of the e2e leg, on the
        # socket this GPU hangs off.  Opt-in: the GPU pool's boxes are single-node VMs (nvidia-smi
        # topo: one NUMA node, 32 CPUs), where there is nothing to choose.
        from ggs_b200.numa import bind_to_device
        numa = bind_to_device(local_rank)
        log(f"[bench] rank {rank}: NUMA node {numa['node']}, bound to {numa['cpus']} CPUs: {numa['bound']}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    H, W, N, P = wl["H"], wl["W"], wl["N"], wl["P"]
    K, Wm = args.steps, args.warmup
    t_np = synth.synthetic_target_np(H, W, 0)
    target = torch.from_numpy(t_np).to(dev)
    # the weight mask of algorithm.py:42-49, computed on the device by the library
    mask = ggs_b200.importance_mask(target, H, W, edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3,
                                    gamma=0.7, floor=0.15, smooth=3, strength=0.7) if wl["mask"] else None
    m_np = None if mask is None else mask.cpu().numpy()

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

    # N > 1: the fitness vector reaches every rank through peer-to-peer stores issued by the raster
    # kernel itself (ggs_b200.peers: no collective launch after the evaluation); --gather nccl,
    # or GPUs that cannot map each other's memory, use one NCCL all-gather instead.
    all_fit = torch.empty((world * P,), dtype=torch.float32, device=dev) if world > 1 else None
    peers, gather = None, ("none" if world == 1 else "nccl")
    if world > 1 and args.gather == "p2p":
        ok_flag = torch.ones(1, device=dev)
        try:
            from ggs_b200.peers import PeerGroup
            peers = PeerGroup.from_process_group(capacity=max(world * P, 8192), device=dev)
        except Exception as e:
            log(f"[bench rank {rank}] peer-to-peer exchange unavailable: {e}")
            ok_flag.zero_()
        dist.all_reduce(ok_flag, op=dist.ReduceOp.MIN)       # all ranks or none
        if float(ok_flag) == 0.0:
            peers = None
        gather = "p2p" if peers is not None else "nccl"

    def step(i):
        if peers is not None:
            return peers.fitness_allgather(pool[i % POOL], target, H, W, offset=rank * P,
                                           total=world * P, weight_mask=mask)
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
    last = last.clone()          # a peer-gathered vector is a view that lives for two gathers
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

    # ---- N > 1: the gathered vector must equal a single-GPU evaluation of the same candidates,
    # bit for bit (outside the timed region): rank 0 regenerates every rank's shard of the last
    # step and evaluates it alone.
    gather_ok = None
    if world > 1:
        r_last = (Wm + K - 1) % POOL
        if rank == 0:
            whole = torch.cat([torch.from_numpy(synth.new_population_np(P, N, H, W, seed=42 + 1000 * q + r_last))
                               for q in range(world)]).to(dev)
            alone = ggs_b200.fitness(whole, target, H, W, 3.0, weight_mask=mask)
            gather_ok = bool(torch.equal(alone, last))
            del whole
            assert gather_ok, "gathered fitness vector differs from the single-GPU evaluation"
        dist.barrier()

    # ---- N == 1: the reference's real (Triton) path on this GPU, live, and parity against it
    ref_gpu = None
    if world == 1 and not args.no_reference_gpu:
        ref_gpu, ref_npz = reference_gpu_live(args, wl, local_rank)
        if ref_npz and os.path.exists(ref_npz):
            z = np.load(ref_npz)
            rm = torch.from_numpy(z["mask"]).to(dev) if wl["mask"] else None
            ours = ggs_b200.fitness(pool[0], target, H, W, 3.0, weight_mask=rm).double().cpu().numpy()
            theirs = z["fitness"]
            ra, rb = np.argsort(theirs, kind="stable"), np.argsort(ours, kind="stable")
            ref_gpu["parity_vs_this_library"] = {
                "candidates": int(P), "fitness_max_rel_diff": float(np.abs(ours / theirs - 1.0).max()),
                "ranking_positions_differing": int((ra != rb).sum()),
                "top8_identical": bool(np.array_equal(ra[:8], rb[:8])),
                "note": "same genomes (seed 42), same target, the reference's own mask"}
            ref_gpu["speedup_e2e_list_api"] = None  # filled below once e2e is known

    # ---- BASELINE config 4 as stated: 512x512, 4,000 splats, population 8,192 FIXED, split over N
    c4 = None
    if not args.no_config4 and args.workload == "c3" and not (args.side or args.splats or args.population):
        try:
            c4 = config4_leg(rank, world, dev, peers)
        except Exception as e:
            c4 = {"error": f"{type(e).__name__}: {e}"[:300]}
            log(f"[bench rank {rank}] config 4 leg failed: {e}")

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
        peak = float(peaks["ffma_tflops"])
        # flops the kernel really spends per EVALUATED pixel-splat pair:
        # recurrence path 100 per 8 pixels (15 FFMA2, 9 FMUL2, 4 FADD2, 8 scalar FP, 4 MUFU),
        # exact path 2 Horner FFMA2 + 2 MUFU + 5 blend operations per pixel pair + the set-up
        real_flop = 12.5 * work["recurrence_pairs"] + 17.0 * work["exact_pairs"]
        roofline = {
            "bound": "fp32", "kernel": "ggs::raster_kernel", "achieved": achieved,
            "peak": peak, "unit": "TFLOP/s",
            "frac": None if achieved is None else achieved / peak,
            "traffic": ncu_traffic(args.workload, P),
            "peak_source": "scalar FFMA rate measured on this GPU by ggs_probe_peaks in this run "
                           "(MEASURED_PEAKS.json has HBM and bf16 tensor entries only, no fp32); "
                           "nominal = 148 SM x 128 lanes x 2 x clocks.max.sm 1965 MHz",
            "nominal_fp32_tflops": NOMINAL_FP32_TFLOPS,
            "frac_of_nominal": None if achieved is None else achieved / NOMINAL_FP32_TFLOPS,
            "measured_ffma_tflops": peaks["ffma_tflops"],
            "measured_ffma2_tflops": peaks["ffma2_tflops"],
            "measured_mufu_ex2_gops": peaks["mufu_ex2_gops"],
            "evaluated_pairs": work["pairs"],
            "frac_evaluated": (real_flop / (kt["raster_ms"] / max(1, kt["evaluations"]) * 1e-3) / 1e12 / peak
                               if kt["raster_ms"] > 0 else None),
            "frac_evaluated_note": "flops the kernel really executes on the pairs it evaluates "
                                   "(12.5 per pair on the recurrence path, 17 on the exact path) / "
                                   "raster time / peak: survives the saturation stop, unlike `frac`, "
                                   "which credits the reference's 23 flops for every in-AABB pair",
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
        if ref_gpu and "value" in ref_gpu:
            ref_gpu["speedup_e2e_list_api"] = e2e_value / ref_gpu["value"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_all / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl, world),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": P * N * 9 * 4, "d2h_bytes_per_step": P * 4,
                    "ms_per_step": e2e_ms_all / K,
                    "api": "ggs_ctx_fitness_host (C ABI, pinned host genomes in, host fitness out)"},
            "gpu_launches": (3 if gather == "p2p" else 2) * K,   # decode + raster (+ the one-warp peer wait)
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "reference_gpu": ref_gpu,
            "gather": {"how": {"p2p": "fitness values stored into every rank's vector by the raster kernel "
                                      "over NVLink (CUDA IPC peer memory), flags + one-warp wait; no collective",
                               "nccl": "dist.all_gather_into_tensor after the evaluation",
                               "none": "single GPU"}[gather], "kind": gather},
            "gather_bit_identical": gather_ok,
            "numa": numa,
            "extra": {"config4": c4},
        }
        emit(line)
    if peers is not None:
        peers.check()
        dist.barrier()
        peers.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference", "reference-gpu-child"], default="ours")
    ap.add_argument("--gather", choices=["p2p", "nccl"], default="p2p",
                    help="N > 1: how the fitness vector reaches every rank")
    ap.add_argument("--numa-bind", action="store_true",
                    help="N > 1: pin each rank to the CPUs of its GPU's NUMA node before the pinned host "
                         "buffers of the e2e leg are allocated (multi-socket hosts)")
    ap.add_argument("--no-config4", action="store_true",
                    help="skip the BASELINE config 4 leg (512x512, 4,000 splats, P = 8,192 over N GPUs)")
    ap.add_argument("--no-reference-gpu", action="store_true",
                    help="skip the live run of the reference's Triton path (N = 1)")
    ap.add_argument("--ref-out", default="", help=argparse.SUPPRESS)
    ap.add_argument("--ref-budget", type=float, default=25.0, help=argparse.SUPPRESS)
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
    if args.impl == "reference-gpu-child":
        return run_reference_gpu_child(args, wl)
    return run_reference(args, wl) if args.impl == "reference" else run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
