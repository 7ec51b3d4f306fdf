#!/usr/bin/env python
"""Latency of the fitness entry points at small batch (the SA regime, BASELINE config 2)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import torch
import ggs_b200
from ggs_b200 import synth
from modules.fitness import fitness_many, fitness_population
H = W = 256; N = 500
t_np = synth.synthetic_target_np(H, W, 0)
target = torch.from_numpy(t_np).cuda(); mask = torch.from_numpy(synth.importance_mask_np(t_np)).cuda()
for B in (1, 8, 64):
    g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=1)).cuda()
    pop = list(g.unbind(0))
    def t(fn, n=300):
        for _ in range(20): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B:3d}: device time/call {e0.elapsed_time(e1)*10:.1f} us | async fitness(tensor) {t(lambda: ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask)):.1f} us | "
          f"fitness_many(list) {t(lambda: fitness_many(pop, target, H, W, 3.0, 'cuda', weight_mask=mask)):.1f} us | "
          f"fitness_population(list)->floats {t(lambda: fitness_population(pop, target, H, W, 3.0, 'cuda', weight_mask=mask)):.1f} us")
