#!/usr/bin/env python
"""Device time per GA generation / SA iteration on the engines (ggs_ga_run / ggs_sa_run), CUDA
events around a block of enqueued steps, with programmatic dependent launch on and off
(ggs_set_option("pdl", ...) switches it, so both run in one process on the same state)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import numpy as np
import torch
import ggs_b200
import modules.config as C
from ggs_b200 import synth
from ggs_b200.engine import GaEngine, SaEngine
from modules.utils import build_mut_sigma, scale_log_bounds

def events(fn, reps):
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); n = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best * 1e3   # us per step

def ga(side, N, P, steps):
    H = W = side
    t = synth.synthetic_target_np(H, W, 3)
    tgt, m = torch.from_numpy(t).cuda(), torch.from_numpy(synth.importance_mask_np(t)).cuda()
    pop = torch.from_numpy(synth.new_population_np(P, N, H, W, seed=1)).cuda()
    lo, hi = scale_log_bounds(H, W, C.MIN_SCALE_SPLATS, C.MAX_SCALE_SPLATS)
    total = 8 * steps + 8
    eng = GaEngine(tgt, m, H, W, P, N, min(C.ELITE_K, P), total)
    eng.start(pop, 7)
    rows = [build_mut_sigma(1, 100, C.SCHEDULE, C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN)] * steps
    def block():
        eng.run(rows, C.TOUR_K, C.CXPB, C.MUTPB, lo, hi); return steps
    out = {}
    for pdl in ("1", "0", "1", "0"):
        ggs_b200.set_option("pdl", int(pdl))
        out.setdefault(pdl, []).append(events(block, 2))
    eng.close()
    print(f"GA  {side}x{side}, {N} splats, population {P}: {min(out['1']):9.1f} us/generation with PDL, "
          f"{min(out['0']):9.1f} without ({1e6 / min(out['1']):.0f} generations/s)")

def sa(side, N, tries, steps, batched=True):
    H = W = side
    t = synth.synthetic_target_np(H, W, 3)
    tgt, m = torch.from_numpy(t).cuda(), torch.from_numpy(synth.importance_mask_np(t)).cuda()
    state = torch.from_numpy(synth.new_population_np(1, N, H, W, seed=1)).cuda()[0]
    lo, hi = scale_log_bounds(H, W, C.MIN_SCALE_SPLATS, C.MAX_SCALE_SPLATS)
    eng = SaEngine(tgt, m, H, W, N, tries, 8 * steps + 8, batch_neighbors=batched)
    eng.start(state, 7)
    rows = [build_mut_sigma(1, 100, C.SCHEDULE, C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN)] * steps
    uni = np.random.default_rng(0).random((steps, tries))
    def block():
        eng.run(rows, [1e-4] * steps, uni, C.MUTPB, lo, hi); return steps
    out = {}
    for pdl in ("1", "0", "1", "0"):
        ggs_b200.set_option("pdl", int(pdl))
        out.setdefault(pdl, []).append(events(block, 2))
    eng.close()
    print(f"SA  {side}x{side}, {N} splats, {tries} tries ({'batched' if batched else 'sequential, the reference chain'}): "
          f"{min(out['1']):9.1f} us/iteration with PDL, "
          f"{min(out['0']):9.1f} without ({1e6 / min(out['1']):.0f} iterations/s)")

ga(128, 100, 32, 400)
ga(256, 512, 32, 400)
ga(256, 1000, 1024, 30)
ga(512, 4000, 1024, 6)
sa(256, 500, 8, 400)
sa(256, 500, 1, 400)
sa(256, 500, 64, 200)
sa(256, 500, 8, 200, batched=False)
sa(256, 512, 8, 200, batched=False)
