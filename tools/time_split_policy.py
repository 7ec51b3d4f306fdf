#!/usr/bin/env python
"""Does ggs_choose_split still pick the fastest split?  Device time per evaluation (CUDA-graph
replays) for every split at small batch sizes, next to the automatic choice."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200"), os.path.join(ROOT, "tools")]
import torch
import ggs_b200
from ggs_b200 import synth
from time_cta_order import device_us  # noqa: E402

for side, N in ((256, 500), (128, 100), (512, 1000)):
    for B in (1, 2, 3, 4, 6, 8, 12, 16):
        H = W = side
        t_np = synth.synthetic_target_np(H, W, 0)
        target = torch.from_numpy(t_np).cuda()
        mask = torch.from_numpy(synth.importance_mask_np(t_np)).cuda()
        g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=1)).cuda()
        t = {}
        for s in (1, 2, 4, 8):
            t[s] = device_us(lambda: ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask, split=s), n=20, reps=5)
        auto = ggs_b200.choose_split(B, N, H, W)
        best = min(t, key=t.get)
        flag = "" if t[auto] <= 1.03 * t[best] else f"   <-- split {best} is {100 * (t[auto] / t[best] - 1):.0f} % faster"
        print(f"{side}x{side}, {N} splats, B {B:2d}: " + "  ".join(f"split {s}: {t[s]:6.1f}" for s in t) +
              f"   automatic {auto}{flag}", flush=True)
