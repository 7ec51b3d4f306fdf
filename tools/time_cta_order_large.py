#!/usr/bin/env python
"""Where does the centre-out CTA order stop paying?  Grids of 4 to 55 waves, candidate-major
(tile_order 0) against the default (tile_order 1: centre-out up to sixteen waves)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200"), os.path.join(ROOT, "tools")]
import torch
import ggs_b200
from ggs_b200 import synth


def device_us(fn, n=10, reps=6):
    for _ in range(3):
        fn()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


for side, N, B in ((256, 1000, 80), (256, 1000, 128), (256, 1000, 256), (256, 1000, 1024), (256, 500, 128),
                   (128, 100, 512), (128, 1000, 1024), (512, 1000, 32), (512, 1000, 64)):
    H = W = side
    t_np = synth.synthetic_target_np(H, W, 0)
    target = torch.from_numpy(t_np).cuda()
    mask = torch.from_numpy(synth.importance_mask_np(t_np)).cuda()
    g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=1)).cuda()
    r = []
    for order in (0, 1, 0, 1):
        ggs_b200.set_option("tile_order", order)
        r.append(device_us(lambda: ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask)))
    ctas = B * ((side + 31) // 32) ** 2
    print(f"{side}x{side}, {N} splats, {B} candidates ({ctas} CTAs, {ctas / 1184:.1f} waves): candidate-major "
          f"{min(r[0], r[2]):.1f} us, default order {min(r[1], r[3]):.1f} us", flush=True)
ggs_b200.set_option("tile_order", 1)
