#!/usr/bin/env python
"""One small evaluation per listed variant, for `ncu --set full -k regex:raster`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import torch
import ggs_b200
from ggs_b200 import synth
ggs_b200.set_option("fuse", 0)
import os
CASES = {"c1": (128, 100, 32, 1), "c2": (256, 500, 8, 2), "b1": (256, 500, 1, 8)}
for side, N, B, split in [CASES[os.environ.get("CASE", "c1")]]:
    H = W = side
    target = torch.from_numpy(synth.synthetic_target_np(H, W, 0)).cuda()
    g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=1)).cuda()
    for _ in range(3):
        ggs_b200.fitness(g, target, H, W, 3.0, split=split)
    torch.cuda.synchronize()
