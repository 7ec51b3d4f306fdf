#!/usr/bin/env python
"""Centre-out CTA order on grids of LESS than one wave (several CTAs per SM, all resident at once):
the order in which the block scheduler deals the tiles to the SMs matters -- the SMs that get one
CTA more than the others should get a cheap one.  Prints tools/time_cta_order.py's table first."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200"), os.path.join(ROOT, "tools")]
import torch
import ggs_b200
from ggs_b200 import synth
from time_cta_order import device_us  # noqa: E402  (runs that tool's table first)

for side, N, B, split in ((128, 100, 32, 1), (256, 500, 8, 1), (256, 500, 16, 1), (256, 512, 12, 1), (128, 100, 64, 1),
                          (512, 1000, 4, 1), (256, 1000, 8, 1), (256, 500, 1, 8), (256, 500, 2, 4), (256, 500, 4, 2),
                          (512, 1000, 1, 2), (128, 100, 16, 2)):
    H = W = side
    t_np = synth.synthetic_target_np(H, W, 0)
    target = torch.from_numpy(t_np).cuda()
    mask = torch.from_numpy(synth.importance_mask_np(t_np)).cuda()
    g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=1)).cuda()
    r = []
    for order in (0, 1, 0, 1):
        ggs_b200.set_option("tile_order", order)
        r.append(device_us(lambda: ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask, split=split)))
    ctas = B * ((side + 31) // 32) ** 2 * split
    print(f"{side}x{side}, {N} splats, {B} candidates ({ctas} CTAs, split {split}): candidate-major {min(r[0], r[2]):.1f} us, "
          f"centre-out {min(r[1], r[3]):.1f} us", flush=True)
ggs_b200.set_option("tile_order", 1)
