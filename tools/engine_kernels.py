#!/usr/bin/env python
"""A few engine steps for `ncu --metrics gpu__time_duration.sum` (per-kernel durations of a GA
generation / an SA iteration): SA batched 8 tries, SA sequential 8 tries, GA at the reference's
default sizes, GA at config 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import numpy as np
import torch
import modules.config as C
from ggs_b200 import synth
from ggs_b200.engine import GaEngine, SaEngine
from modules.utils import build_mut_sigma, scale_log_bounds

def inputs(side, N, P):
    t = synth.synthetic_target_np(side, side, 3)
    tgt = torch.from_numpy(t).cuda()
    m = torch.from_numpy(synth.importance_mask_np(t)).cuda()
    return tgt, m, torch.from_numpy(synth.new_population_np(P, N, side, side, seed=1)).cuda()

rows = [build_mut_sigma(1, 100, C.SCHEDULE, C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN)] * 3
for batched in (True, False):
    tgt, m, pop = inputs(256, 500, 1)
    lo, hi = scale_log_bounds(256, 256, 3.0, 0.1)
    eng = SaEngine(tgt, m, 256, 256, 500, 8, 16, batch_neighbors=batched)
    eng.start(pop[0], 7)
    eng.run(rows, [1e-4] * 3, np.random.default_rng(0).random((3, 8)), C.MUTPB, lo, hi)
    eng.state()
    eng.close()
    print("SA", "batched" if batched else "sequential", flush=True)
for side, N, P in ((256, 512, 32), (128, 100, 32)):
    tgt, m, pop = inputs(side, N, P)
    lo, hi = scale_log_bounds(side, side, 3.0, 0.1)
    eng = GaEngine(tgt, m, side, side, P, N, 8, 16)
    eng.start(pop, 7)
    eng.run(rows, C.TOUR_K, C.CXPB, C.MUTPB, lo, hi)
    eng.state()
    eng.close()
    print("GA", side, N, P, flush=True)
