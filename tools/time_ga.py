#!/usr/bin/env python
"""Where does a GA generation spend its time?  (BASELINE config 3 shape by default.)
    P=1024 N=1000 python tools/time_ga.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
import torch
import modules.config as C
from modules.fitness import fitness_many
from modules.genetic import breed_population
from oracle.torch_ref import crossover_population, mutate_population, tournament_indices
from modules.population import new_population
from ggs_b200 import synth

P, N, H, W = int(os.environ.get("P", 1024)), int(os.environ.get("N", 1000)), 256, 256
dev = "cuda"
t_np = synth.synthetic_target_np(H, W, 0)
target = torch.from_numpy(t_np).to(dev); mask = torch.from_numpy(synth.importance_mask_np(t_np)).to(dev)
pop = new_population(P, N, H, W, 3.0, 0.1, device=dev)
fit = fitness_many(pop, target, H, W, 3.0, dev, weight_mask=mask)

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, out

args = (5, 100, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN)
ms_breed, off = timed(lambda: breed_population(pop, fit, *args, C.TOUR_K, C.CXPB, C.MUTPB, H, W, 3.0, 0.1, seed=1))
def torch_ops():
    parents = pop[tournament_indices(fit, P, 2)][torch.randperm(P, device=dev)]
    o = crossover_population(parents, C.CXPB)
    return mutate_population(o, *args, C.MUTPB, H, W, 3.0, 0.1)
ms_torch, _ = timed(torch_ops)
ms_fit, f2 = timed(lambda: fitness_many(off, target, H, W, 3.0, dev, weight_mask=mask))
def elit():
    idx = torch.argsort(fit, stable=True)[:8]
    return torch.cat([pop[idx], off[:P - 8]]), torch.cat([fit[idx], f2[:P - 8]]).cpu().tolist()
ms_el, _ = timed(elit)
print(f"P={P} N={N} {H}x{W}: breed kernel {ms_breed:.3f} ms (same operators as batched torch ops: "
      f"{ms_torch:.3f} ms), fitness {ms_fit:.3f}, elitism + host copy {ms_el:.3f} -> "
      f"{ms_breed + ms_fit + ms_el:.3f} ms/generation")
