#!/usr/bin/env python
"""Breeding-kernel time and bandwidth at a given population shape.
    P=8192 N=4000 SIDE=512 python tools/time_breed.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import torch
import modules.config as C
from modules.genetic import breed_population
from modules.population import new_population

P, N, S = int(os.environ.get("P", 8192)), int(os.environ.get("N", 4000)), int(os.environ.get("SIDE", 512))
pop = new_population(P, N, S, S, 3.0, 0.1, device="cuda")
fit = torch.rand(P, device="cuda")
out = torch.empty_like(pop)
def run(): return breed_population(pop, fit, 5, 100, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN,
                                   C.TOUR_K, C.CXPB, C.MUTPB, S, S, 3.0, 0.1, seed=1, out=out)
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
gb = 2 * P * N * 36 / 1e9
print(f"breed P={P} N={N}: {ms:.3f} ms, {gb / ms * 1e3:.0f} GB/s (read parents + write children = {gb:.2f} GB)")
def misc():
    order = torch.argsort(fit, stable=True); r = fit[order].double()
    s = torch.stack([r[0], r.mean(), r[P // 2]]).tolist(); out[:8] = pop[order[:8]]
    return s
for _ in range(3): misc()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): misc()
torch.cuda.synchronize(); print(f"sort + stats + elite copy: {(time.perf_counter() - t0) * 100:.3f} ms")
