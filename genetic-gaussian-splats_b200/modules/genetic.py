"""GA operators (reference: modules/genetic.py:8-92).

  * tournament_selection / crossover_uniform keep the reference's per-individual signatures
    for callers that hold a list of [N,9] tensors (host logic over torch, any device);
  * mutate_individual (same signature, in place) and breed_population (the whole
    selection -> crossover -> mutation step of algorithm.py:87-120 over the resident [P,N,9]
    tensor) are ONE launch of the library's breeding kernel (ggs_ga_breed): Philox
    counter-based streams, no per-individual Python, no .item() syncs.  There is no CPU path;
    the batched torch operators that serve as the kernel's reference live with the tests
    (oracle/torch_ref.py).  Same distributions as the reference, different random streams
    (GA trajectory parity is not a goal).
"""
import random
from typing import List

import torch

from modules.population import duplicate_individual
from modules.utils import build_mut_sigma, scale_log_bounds


def tournament_selection(pop: List[torch.Tensor], fits: List[float], k: int = 2) -> torch.Tensor:
    """Best (lowest fitness) of k uniformly drawn individuals; returns a copy."""
    picks = [random.randrange(len(pop)) for _ in range(k)]
    return duplicate_individual(pop[min(picks, key=lambda i: fits[i])])


def crossover_uniform(a: torch.Tensor, b: torch.Tensor, p: float = 0.5):
    """Swap whole splats (rows) between two parents with probability p per row."""
    take_a = torch.rand((a.shape[0], 1), device=a.device) < p
    return torch.where(take_a, a, b), torch.where(take_a, b, a)


def _draw_seed() -> int:
    """A fresh 63-bit key for the kernel's counter-based streams; follows torch.manual_seed."""
    hi, lo = torch.randint(0, 2**31 - 1, (2,)).tolist()
    return (hi << 31) | lo


@torch.no_grad()
def mutate_individual(ind: torch.Tensor, is_elite: bool, gen: int, total_gens: int,
                      schedule: str, mut_sigma_max: dict, mut_sigma_min: dict,
                      mutpb: float, H: int, W: int, min_scale_splats: float,
                      max_scale_splats: float):
    """In-place Gaussian mutation of one [N,9] individual (genetic.py:32-92); `is_elite` is
    unused, as in the reference.  One launch of the breeding kernel with a one-individual
    population and no crossover -- exactly one mutate_individual call."""
    from ggs_b200 import breed
    assert torch.cuda.is_available(), "mutate_individual needs a CUDA device (no CPU path)"
    src = ind if ind.is_cuda else ind.cuda()
    lo, hi = scale_log_bounds(H, W, min_scale_splats, max_scale_splats)
    sigma = build_mut_sigma(gen, total_gens, schedule, mut_sigma_max, mut_sigma_min)
    out = breed(src.unsqueeze(0), torch.zeros(1, device=src.device), sigma, tour_k=1, cxpb=0.0,
                mutpb=mutpb, log_scale_lo=lo, log_scale_hi=hi, seed=_draw_seed(), generation=gen)
    ind[..., :9] = out[0].to(device=ind.device, dtype=ind.dtype)
    return ind


@torch.no_grad()
def breed_population(pop: torch.Tensor, fitness: torch.Tensor, gen: int, total_gens: int,
                     schedule: str, mut_sigma_max: dict, mut_sigma_min: dict, tour_k: int,
                     cxpb: float, mutpb: float, H: int, W: int, min_scale_splats: float,
                     max_scale_splats: float, seed: int = 0, out=None) -> torch.Tensor:
    """Selection + crossover + mutation of the whole population -> offspring [P,N,9]
    (written into `out` when given: a contiguous [P,N,9] tensor that does not alias `pop`).
    ONE kernel launch (ggs_ga_breed in libggs_b200.so, Philox counter-based randomness keyed
    by (seed, gen))."""
    from ggs_b200 import breed
    assert pop.is_cuda, "breed_population needs the population on a CUDA device (no CPU path)"
    lo, hi = scale_log_bounds(H, W, min_scale_splats, max_scale_splats)
    sigma = build_mut_sigma(gen, total_gens, schedule, mut_sigma_max, mut_sigma_min)
    return breed(pop, fitness, sigma, tour_k=tour_k, cxpb=cxpb, mutpb=mutpb, log_scale_lo=lo,
                 log_scale_hi=hi, seed=seed, generation=gen, out=out)
