#!/usr/bin/env python
"""Candidate-major against interior-first CTA order (ggs_set_option("tile_order", ...)) on grids
of one to four waves: device time per evaluation from CUDA-graph replays."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200"), os.path.join(ROOT, "tools")]
import torch
import ggs_b200
from ggs_b200 import synth


def device_us(fn, n=40, reps=8):
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


for side, N, B in ((256, 512, 24), (256, 512, 32), (256, 1000, 40), (128, 100, 128), (512, 1000, 12),
                   (256, 500, 64), (512, 4000, 16)):
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
    print(f"{side}x{side}, {N} splats, {B} candidates ({ctas} CTAs): candidate-major {min(r[0], r[2]):.1f} us, "
          f"interior first {min(r[1], r[3]):.1f} us", flush=True)
ggs_b200.set_option("tile_order", 1)
