#!/usr/bin/env python
"""One config-2 / config-1 evaluation per CTA order, for
   ncu --metrics sm__inst_executed.avg,sm__inst_executed.max,sm__inst_executed.min,sm__cycles_active.avg,sm__cycles_active.max,gpu__time_duration.sum -k regex:raster
how evenly is the work spread over the SMs?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import torch
import ggs_b200
from ggs_b200 import synth
CASES = {"c1": (128, 100, 32), "c2": (256, 500, 8), "ga": (256, 512, 24)}
side, N, B = CASES[os.environ.get("CASE", "c2")]
H = W = side
t_np = synth.synthetic_target_np(H, W, 0)
target = torch.from_numpy(t_np).cuda()
mask = torch.from_numpy(synth.importance_mask_np(t_np)).cuda()
g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=1)).cuda()
for order in (0, 1):
    ggs_b200.set_option("tile_order", order)
    for _ in range(3):
        ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask, split=1)
    torch.cuda.synchronize()
