"""GA operators (reference: modules/genetic.py:8-92).

Two forms of every operator:
  * the reference's per-individual API (tournament_selection, crossover_uniform,
    mutate_individual) with identical signatures, for callers that hold a list of [N,9] tensors;
  * batched forms over the resident [P,N,9] population tensor (tournament_indices,
    crossover_population, mutate_population): no per-individual Python, no .item() syncs.  These
    are what modules/algorithm.py uses.  Both draw from the same distributions; the random
    streams differ (GA trajectory parity is not a goal).
"""
import random
from typing import List

import torch

from modules.population import duplicate_individual
from modules.utils import build_mut_sigma, clamp_genome, wrap_angle


# ------------------------------------------------------------------ per-individual (reference API)

def tournament_selection(pop: List[torch.Tensor], fits: List[float], k: int = 2) -> torch.Tensor:
    """Best (lowest fitness) of k uniformly drawn individuals; returns a copy."""
    picks = [random.randrange(len(pop)) for _ in range(k)]
    return duplicate_individual(pop[min(picks, key=lambda i: fits[i])])


def crossover_uniform(a: torch.Tensor, b: torch.Tensor, p: float = 0.5):
    """Swap whole splats (rows) between two parents with probability p per row."""
    take_a = torch.rand((a.shape[0], 1), device=a.device) < p
    return torch.where(take_a, a, b), torch.where(take_a, b, a)


def _ensure_one_true(mask: torch.Tensor) -> torch.Tensor:
    """At least one True in the whole mask (in place)."""
    if not mask.any():
        mask.view(-1)[int(torch.randint(mask.numel(), (1,), device=mask.device))] = True
    return mask


def mutate_individual(ind: torch.Tensor, is_elite: bool, gen: int, total_gens: int,
                      schedule: str, mut_sigma_max: dict, mut_sigma_min: dict,
                      mutpb: float, H: int, W: int, min_scale_splats: float,
                      max_scale_splats: float):
    """In-place Gaussian mutation of one [N,9] individual; `is_elite` is unused (as in the
    reference).  Implemented on the batched operator with P = 1."""
    mutate_population(ind.unsqueeze(0), gen, total_gens, schedule, mut_sigma_max, mut_sigma_min,
                      mutpb, H, W, min_scale_splats, max_scale_splats)
    return ind


# ----------------------------------------------------------------------------- batched, on device

@torch.no_grad()
def breed_population(pop: torch.Tensor, fitness: torch.Tensor, gen: int, total_gens: int,
                     schedule: str, mut_sigma_max: dict, mut_sigma_min: dict, tour_k: int,
                     cxpb: float, mutpb: float, H: int, W: int, min_scale_splats: float,
                     max_scale_splats: float, seed: int = 0, out=None) -> torch.Tensor:
    """Selection + crossover + mutation of the whole population -> offspring [P,N,9]
    (written into `out` when given: a contiguous [P,N,9] tensor that does not alias `pop`).

    On a CUDA population this is ONE kernel launch (ggs_ga_breed in libggs_b200.so, Philox
    counter-based randomness keyed by (seed, gen)); on a CPU population (tests) it is the
    composition of the batched torch operators below.  Same operators either way."""
    from modules.utils import scale_log_bounds
    if pop.is_cuda:
        from ggs_b200 import breed
        lo, hi = scale_log_bounds(H, W, min_scale_splats, max_scale_splats)
        sigma = build_mut_sigma(gen, total_gens, schedule, mut_sigma_max, mut_sigma_min)
        return breed(pop, fitness, sigma, tour_k=tour_k, cxpb=cxpb, mutpb=mutpb, log_scale_lo=lo,
                     log_scale_hi=hi, seed=seed, generation=gen, out=out)
    P = pop.shape[0]
    parents = pop[tournament_indices(fitness, P, k=tour_k)]
    parents = parents[torch.randperm(P, device=pop.device)]
    offspring = crossover_population(parents[..., :9].contiguous(), cxpb)
    offspring = mutate_population(offspring, gen, total_gens, schedule, mut_sigma_max,
                                  mut_sigma_min, mutpb, H, W, min_scale_splats, max_scale_splats)
    if out is not None:
        out.copy_(offspring)
        return out
    return offspring


@torch.no_grad()
def tournament_indices(fitness: torch.Tensor, n_parents: int, k: int = 2,
                       generator=None) -> torch.Tensor:
    """[n_parents] indices: each the best of k uniform draws (genetic.py:8-14, batched)."""
    P = fitness.shape[0]
    draws = torch.randint(P, (n_parents, k), device=fitness.device, generator=generator)
    best = fitness[draws].argmin(dim=1, keepdim=True)
    return draws.gather(1, best).squeeze(1)


@torch.no_grad()
def crossover_population(parents: torch.Tensor, cxpb: float, p: float = 0.5,
                         generator=None) -> torch.Tensor:
    """Pairs (2i, 2i+1) of the (already shuffled) parents exchange rows with probability p
    when the pair is selected for crossover (probability cxpb); otherwise both are copied
    (algorithm.py:94-100 + genetic.py:17-21, batched).  An odd last parent is copied."""
    P, N, _ = parents.shape
    out = parents.clone()
    npairs = P // 2
    if npairs == 0:
        return out
    a, b = parents[0:2 * npairs:2], parents[1:2 * npairs:2]
    dev = parents.device
    do_cx = torch.rand((npairs, 1, 1), device=dev, generator=generator) < cxpb
    take_a = torch.rand((npairs, N, 1), device=dev, generator=generator) < p
    keep = take_a | ~do_cx
    out[0:2 * npairs:2] = torch.where(keep, a, b)
    out[1:2 * npairs:2] = torch.where(keep, b, a)
    return out


def _ensure_one_true_rows(mask: torch.Tensor, generator=None) -> torch.Tensor:
    """Every individual (dim 0) gets at least one True somewhere in its mask."""
    P = mask.shape[0]
    flat = mask.reshape(P, -1)
    empty = ~flat.any(dim=1)
    pick = torch.randint(flat.shape[1], (P,), device=mask.device, generator=generator)
    flat[torch.arange(P, device=mask.device)[empty], pick[empty]] = True
    return flat.reshape(mask.shape)


@torch.no_grad()
def mutate_population(pop: torch.Tensor, gen: int, total_gens: int, schedule: str,
                      mut_sigma_max: dict, mut_sigma_min: dict, mutpb: float, H: int, W: int,
                      min_scale_splats: float, max_scale_splats: float, generator=None):
    """In-place mutation of every individual of pop [P,N,9] (genetic.py:32-92, batched):
    per-gene Bernoulli(mutpb) masks for xy / log-scales / theta / rgb / alpha with at least one
    mutated gene per group and individual, annealed Gaussian noise, projection onto the legal
    box, then one "bring a bigger splat forward" swap per individual."""
    SIG = build_mut_sigma(gen, total_gens, schedule, mut_sigma_max, mut_sigma_min)
    P, N, _ = pop.shape
    dev, dt = pop.device, pop.dtype

    def bern(cols):
        return torch.rand((P, N, cols), device=dev, generator=generator) < mutpb

    def noise(cols):
        return torch.randn((P, N, cols), device=dev, dtype=dt, generator=generator)

    m_xy = _ensure_one_true_rows(bern(2), generator)
    m_ab = _ensure_one_true_rows(bern(2), generator)
    m_t = _ensure_one_true_rows(bern(1), generator)
    m_col = _ensure_one_true_rows(bern(2), generator)          # [rgb flag, alpha flag]
    m_rgba = torch.cat([m_col[..., 0:1].expand(-1, -1, 3), m_col[..., 1:2]], dim=-1)

    pop[..., 0:2] += noise(2) * SIG["xy"] * m_xy
    pop[..., 2:4] += noise(2) * torch.tensor([SIG["alog"], SIG["blog"]], device=dev, dtype=dt) * m_ab
    pop[..., 4:5] += noise(1) * SIG["theta"] * m_t
    pop[..., 4] = wrap_angle(pop[..., 4])
    pop[..., 5:9] += noise(4) * torch.tensor([SIG["rgb"]] * 3 + [SIG["alpha"]], device=dev, dtype=dt) * m_rgba
    clamp_genome(pop, H, W, min_scale_splats, max_scale_splats)

    if N >= 2:
        # pick i uniformly in [0, N-2]; among the later splats that are bigger (sigma_x*sigma_y)
        # pick one uniformly and swap it with i (a bigger splat moves towards the back layer)
        rows = torch.arange(P, device=dev)
        i = torch.randint(0, N - 1, (P,), device=dev, generator=generator)
        size = (pop[..., 2] + pop[..., 3]).exp()                       # [P,N]
        later = torch.arange(N, device=dev).unsqueeze(0) > i.unsqueeze(1)
        cand = later & (size > size[rows, i].unsqueeze(1))
        score = torch.rand((P, N), device=dev, generator=generator).masked_fill(~cand, -1.0)
        j = score.argmax(dim=1)
        sel = rows[cand.any(dim=1)]
        if sel.numel() > 0:
            i_s, j_s = i[sel], j[sel]
            tmp = pop[sel, i_s].clone()
            pop[sel, i_s] = pop[sel, j_s]
            pop[sel, j_s] = tmp
    return pop
