#!/usr/bin/env python
"""Run under `ncu --metrics gpu__time_duration.sum --clock-control none --csv`: two evaluations per
(shape, split, fuse) so the per-kernel durations of the small-batch variants can be read off the
launch list (order of launches = order of the loops below)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import torch
import ggs_b200
from ggs_b200 import synth

SHAPES = [(128, 100, 32), (256, 500, 8), (256, 500, 1)]
for side, N, B in SHAPES:
    H = W = side
    target = torch.from_numpy(synth.synthetic_target_np(H, W, 0)).cuda()
    g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=1)).cuda()
    for split in (1, 2, 4, 8):
        for fuse in (0, 1):
            ggs_b200.set_option("fuse", fuse)
            for _ in range(2):
                ggs_b200.fitness(g, target, H, W, 3.0, split=split)
            torch.cuda.synchronize()
            print(f"side {side} N {N} B {B} split {split} fuse {fuse}", flush=True)
