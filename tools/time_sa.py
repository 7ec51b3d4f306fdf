#!/usr/bin/env python
"""Simulated-annealing iterations per second at BASELINE config 2 (256x256, 500 splats, batched
neighbour proposals): the device engine (ggs_sa_run) against the same chain driven from Python.
    N=500 TRIES=8 ITERS=4000 python tools/time_sa.py"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
os.environ.setdefault("TQDM_DISABLE", "1")
import torch
import modules.config as C
from ggs_b200 import synth
from modules.annealing import simulated_annealing

H = W = int(os.environ.get("SIDE", 256))
N, TRIES, ITERS = int(os.environ.get("N", 500)), int(os.environ.get("TRIES", 8)), int(os.environ.get("ITERS", 4000))
target = torch.from_numpy(synth.synthetic_target_np(H, W, 3))
kw = dict(H=H, W=W, device="cuda", n_splats=N, mutpb=C.MUTPB, mut_sigma_max=C.MUT_SIGMA_MAX,
          mut_sigma_min=C.MUT_SIGMA_MIN, sigma_schedule="cosine", min_scale_splats=3.0,
          max_scale_splats=0.1, k_sigma=3.0, mask_strength=0.7, boost_only=False, temp0=1e-4,
          temp_schedule="exp", tries_per_iter=TRIES)
simulated_annealing(target, iterations=50, **kw)   # warm-up (context, mask workspace)
for batched in (False, True):
  kw["batch_neighbors"] = batched
  print("batched neighbours" if batched else "sequential tries (the reference's chain)")
  for loop in ("0", "1"):
      os.environ["GGS_B200_SA_LOOP"] = loop
      zero, long_, energy = [], [], [None, None]
      # alternate 0-iteration and ITERS-iteration runs; the first pair pays one-time costs
      # (allocator growth, lazy module loading) and is dropped
      for rep in range(3):
          for iters in (0, ITERS):
              torch.manual_seed(1); random.seed(1)
              torch.cuda.synchronize(); t0 = time.perf_counter()
              _, e = simulated_annealing(target, iterations=iters, **kw)
              torch.cuda.synchronize(); dt_ = time.perf_counter() - t0
              if rep > 0:
                  (long_ if iters else zero).append(dt_)
              energy[1 if iters else 0] = e
      dt = min(long_) - min(zero)
      print(f"{'python loop' if loop == '1' else 'device engine'}: {H}x{W}, {N} splats, {TRIES} tries/iteration: "
            f"{ITERS} iterations in {dt:.3f} s beyond the {min(zero):.3f} s set-up = {ITERS / dt:.0f} iterations/s "
            f"({ITERS * TRIES / dt:.0f} tries/s); energy {energy[0]:.6f} -> {energy[1]:.6f}")
