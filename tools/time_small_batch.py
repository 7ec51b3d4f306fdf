#!/usr/bin/env python
"""Latency of one evaluation at small batch (BASELINE configs 1 and 2, the SA regime) for every
split / fusion variant of the raster: device time per call from CUDA events around the replay
of a CUDA graph that holds 40 back-to-back calls.  Prints one line per (shape, split, fuse) and the automatic choice."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import torch
import ggs_b200
from ggs_b200 import synth

SHAPES = [("config 1: 128x128, 100 splats, P 32", 128, 100, 32),
          ("config 2: 256x256, 500 splats, 8 neighbours", 256, 500, 8),
          ("256x256, 500 splats, 1 candidate (sequential SA try)", 256, 500, 1),
          ("256x256, 512 splats, 24 children (reference default GA)", 256, 512, 24),
          ("512x512, 4000 splats, 1 candidate (final frame)", 512, 4000, 1)]


def device_us(fn, n=40, reps=8):
    """Device time per call: `n` calls captured in ONE CUDA graph and replayed, so the host's
    per-call overhead (tens of microseconds of Python) does not pace the GPU."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(n):
            fn()
    graph.replay()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


for name, side, N, B in SHAPES:
    H = W = side
    t_np = synth.synthetic_target_np(H, W, 0)
    target = torch.from_numpy(t_np).cuda()
    mask = torch.from_numpy(synth.importance_mask_np(t_np)).cuda()
    g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=1)).cuda()
    auto = ggs_b200.choose_split(B, N, H, W)
    print(f"{name}: automatic split {auto}")
    for split in (1, 2, 4, 8):
        row = []
        for fuse in (0, 1):
            ggs_b200.set_option("fuse", fuse)
            try:
                us = device_us(lambda: ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask, split=split))
                row.append(f"{'fused' if fuse else 'decode+raster'} {us:7.1f} us")
            except ggs_b200.GgsError as e:
                row.append(f"{'fused' if fuse else 'decode+raster'}   n/a")
        print(f"   split {split}: " + " | ".join(row))
    ggs_b200.set_option("fuse", 0)
    us = device_us(lambda: ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask))
    print(f"   default entry: {us:7.1f} us per evaluation, {B / us * 1e6:,.0f} candidates/s")
