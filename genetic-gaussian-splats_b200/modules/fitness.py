"""Drop-in for the reference's modules/fitness.py (fitness.py:8-31 and :35-48).

encode -> decode -> render -> masked squared error -> per-candidate scalar run as one fused
evaluation in libggs_b200.so; candidate images never reach HBM."""
from __future__ import annotations

from typing import List

import torch

from ggs_b200 import LAYOUT_AXES_ANGLE, choose_split
from ggs_b200 import fitness as _fitness


def _stack(pop_batch) -> torch.Tensor:
    if isinstance(pop_batch, torch.Tensor):  # already [B,N,C]: resident population tensor
        return pop_batch
    return torch.stack(list(pop_batch), dim=0)


@torch.no_grad()
def fitness_many(pop_batch: List[torch.Tensor], target: torch.Tensor, H: int, W: int,
                 k_sigma: float, device, tile: int = 32,
                 weight_mask: torch.Tensor | None = None,
                 boost_only: bool = False,
                 boost_beta: float = 1.0):
    """List of axes-angle genomes [N,C>=9] (or one [B,N,C] tensor) -> Tensor[B], lower is
    better.  `tile` is accepted for signature compatibility and ignored."""
    return _evaluate(pop_batch, target, H, W, k_sigma, device, weight_mask, boost_only, boost_beta, 0)


def _evaluate(pop_batch, target, H, W, k_sigma, device, weight_mask, boost_only, boost_beta, split):
    """split: kernel configuration of a larger population this batch is a part of (0 = chosen
    from this batch alone, ggs_choose_split)."""
    G_axes = _stack(pop_batch)
    return _fitness(G_axes, target, int(H), int(W), k_sigma=float(k_sigma),
                    weight_mask=weight_mask, boost_only=bool(boost_only),
                    boost_beta=float(boost_beta), layout=LAYOUT_AXES_ANGLE, device=device,
                    split=split)


@torch.no_grad()
def fitness_population(population: List[torch.Tensor], target: torch.Tensor,
                       H: int, W: int, k_sigma: float, device,
                       tile: int = 32, chunk: int | None = None,
                       weight_mask: torch.Tensor | None = None,
                       boost_only: bool = False) -> List[float]:
    """Python floats, one per individual; `chunk` bounds the candidates per launch."""
    n = len(population)
    if chunk is None or chunk >= n:
        return fitness_many(population, target, H, W, k_sigma, device, tile=tile,
                            weight_mask=weight_mask, boost_only=boost_only).cpu().tolist()
    # chunks are evaluated with the kernel configuration of the WHOLE population, so `chunk`
    # stays what it is in the reference: a memory knob that cannot change a result
    split = choose_split(n, int(population[0].shape[0]), int(H), int(W))
    out: List[float] = []
    for i in range(0, n, chunk):
        out.extend(_evaluate(population[i:i + chunk], target, H, W, k_sigma, device, weight_mask,
                             boost_only, 1.0, split).cpu().tolist())
    return out
