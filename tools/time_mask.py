#!/usr/bin/env python
"""Importance mask: device path (ggs_importance_mask) against the torch ops on the host CPU
(where the reference computes it, algorithm.py:42-49).   python tools/time_mask.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import torch
from ggs_b200 import synth
from modules.mask import compute_importance_mask

kw = dict(edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3, gamma=0.7, floor=0.15, smooth=3, strength=0.7)
for side in (128, 256, 512, 1024, 2048):
    img = torch.from_numpy(synth.synthetic_target_np(side, side, 1))
    dimg = img.cuda()
    for _ in range(3): compute_importance_mask(dimg, side, side, **kw)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): compute_importance_mask(dimg, side, side, **kw)
    torch.cuda.synchronize(); gpu = (time.perf_counter() - t0) / 10
    compute_importance_mask(img, side, side, **kw)
    t0 = time.perf_counter()
    for _ in range(3): compute_importance_mask(img, side, side, **kw)
    cpu = (time.perf_counter() - t0) / 3
    print(f"{side}x{side}: device {gpu * 1e3:.3f} ms, torch on {torch.get_num_threads()} host threads {cpu * 1e3:.2f} ms")
