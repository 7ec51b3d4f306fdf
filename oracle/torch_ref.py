"""Plain-torch (CPU) restatements of two callers of the hot path, kept as the fp32 reference of
their CUDA kernels.  TEST INFRASTRUCTURE ONLY, like everything under oracle/: imported by tests/,
tools/ and bench.py's CPU legs, never by the product package (which has no CPU path).

  importance_mask_torch   compute_importance_mask        /root/reference/modules/mask.py:29-83
  tournament_indices      tournament_selection           /root/reference/modules/genetic.py:8-14
  crossover_population    crossover_uniform + pairing    genetic.py:17-21, algorithm.py:94-100
  mutate_population       mutate_individual              genetic.py:32-92
  breed_population_torch  the selection/crossover/mutation sequence of algorithm.py:87-120

The mask functions follow the reference statement by statement and are pinned by
tests/golden/mask_cases.npz (made by the reference's own mask.py); the genetic operators are
batched over the population and pinned by the distribution fixtures of
tests/golden/breed_reference_stats.npz (sampled from the reference's genetic.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

_LUMA = (0.2126, 0.7152, 0.0722)  # Rec.709


def _unit_range(img: torch.Tensor) -> torch.Tensor:
    return img / 255.0 if img.max() > 1.5 else img


@torch.no_grad()
def _rgb_to_luma(img_hw3: torch.Tensor) -> torch.Tensor:
    """[H,W,3] -> [1,1,H,W] luma."""
    x = _unit_range(img_hw3)
    y = _LUMA[0] * x[..., 0] + _LUMA[1] * x[..., 1] + _LUMA[2] * x[..., 2]
    return y[None, None].contiguous()


def _sobel_edges(y: torch.Tensor) -> torch.Tensor:
    """Gradient magnitude of a [1,1,H,W] map with 3x3 Sobel taps, zero padding."""
    gx_k = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]],
                        dtype=y.dtype, device=y.device).view(1, 1, 3, 3)
    gy_k = gx_k.transpose(2, 3).contiguous()
    gx = F.conv2d(y, gx_k, padding=1)
    gy = F.conv2d(y, gy_k, padding=1)
    return torch.sqrt(gx * gx + gy * gy + 1e-12)


def _local_variance(y: torch.Tensor, k: int = 9) -> torch.Tensor:
    """E[y^2] - E[y]^2 over a k x k box (zero-padded average), floored at 0."""
    half = k // 2
    m1 = F.avg_pool2d(y, k, stride=1, padding=half)
    m2 = F.avg_pool2d(y * y, k, stride=1, padding=half)
    return (m2 - m1 * m1).clamp_min(0)


def _robust01(t: torch.Tensor) -> torch.Tensor:
    """Map the 2nd..98th percentile range to [0,1]."""
    flat = t.flatten()
    lo = torch.quantile(flat, 0.02)
    hi = torch.quantile(flat, 0.98)
    return ((t - lo) / (hi - lo + 1e-12)).clamp(0, 1)


@torch.no_grad()
def importance_mask_torch(
    target_hw3: torch.Tensor, H: int, W: int,
    edge_scales=(1, 2, 4),
    w_edge: float = 0.7,
    w_var: float = 0.3,
    gamma: float = 0.7,
    floor: float = 0.15,
    smooth: int = 0,
    strength: float = 1.0
) -> torch.Tensor:
    """[H0,W0,3] target -> [H,W] weights in [floor', 1]: multi-scale Sobel energy and 9x9
    local variance, each robustly normalised, mixed, optionally box-smoothed, gamma-shaped,
    lifted to `floor` and blended towards 1 by (1 - strength)."""
    x = _unit_range(target_hw3).permute(2, 0, 1)[None]
    x = F.interpolate(x, size=(H, W), mode='bilinear', align_corners=False)
    y = _rgb_to_luma(x[0].permute(1, 2, 0))

    edges = torch.zeros_like(y)
    for s in edge_scales:
        if s > 1:
            e = _sobel_edges(F.avg_pool2d(y, kernel_size=s, stride=s))
            e = F.interpolate(e, size=(H, W), mode='bilinear', align_corners=False)
        else:
            e = _sobel_edges(y)
        edges = edges + e

    mask = _robust01(w_edge * _robust01(edges) + w_var * _robust01(_local_variance(y, k=9)))
    if smooth and smooth > 0:
        mask = _robust01(F.avg_pool2d(mask, kernel_size=smooth, stride=1, padding=smooth // 2))

    mask = (1.0 - floor) * mask.pow(gamma) + floor
    if strength < 1.0:
        mask = (1.0 - strength) * torch.ones_like(mask) + strength * mask
    return mask[0, 0]


def importance_mask_np(target: np.ndarray, strength: float = 0.7) -> np.ndarray:
    """The mask the GA loop builds (algorithm.py:42-49) for a target already at the work size."""
    H, W = target.shape[:2]
    out = importance_mask_torch(torch.from_numpy(np.ascontiguousarray(target, dtype=np.float32)), H, W,
                                edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3, gamma=0.7, floor=0.15,
                                smooth=3, strength=strength)
    return np.ascontiguousarray(out.numpy().astype(np.float32))


# ------------------------------------------------------------------ GA operators, batched torch

def wrap_angle(theta: torch.Tensor) -> torch.Tensor:
    """utils.py:11-12"""
    return (theta + math.pi) % (2 * math.pi) - math.pi


def anneal_factor(gen: int, total: int, kind: str) -> float:
    """utils.py:15-28"""
    g = max(0, min(gen, total))
    p = g / max(1, total)
    if kind == "cosine":
        raw = 0.5 * (1.0 + math.cos(math.pi * p))
    elif kind == "exp":
        raw = (0.2 ** (1.0 / max(1, total))) ** g
    else:  # "linear" and anything unknown
        raw = 1.0 - p
    return max(0.0, raw)


def build_mut_sigma(gen, total, kind, smax: dict, smin: dict) -> dict:
    """utils.py:31-33"""
    f = anneal_factor(gen, total, kind)
    return {k: smin[k] + f * (smax[k] - smin[k]) for k in smax}


def clamp_genome(pop: torch.Tensor, H: int, W: int, min_scale: float, max_scale: float) -> torch.Tensor:
    """utils.py:36-45, over any leading dimensions (in place)."""
    lo, hi = math.log(min_scale), math.log(max_scale * float(max(H, W)))
    pop[..., 0:2].clamp_(0.0, 1.0)
    pop[..., 2:4].clamp_(lo, hi)
    pop[..., 4] = wrap_angle(pop[..., 4])
    pop[..., 5:9].clamp_(0.0, 255.0)
    return pop


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


@torch.no_grad()
def breed_population_torch(pop: torch.Tensor, fitness: torch.Tensor, gen: int, total_gens: int,
                           schedule: str, mut_sigma_max: dict, mut_sigma_min: dict, tour_k: int,
                           cxpb: float, mutpb: float, H: int, W: int, min_scale_splats: float,
                           max_scale_splats: float, generator=None) -> torch.Tensor:
    """Selection -> shuffle -> pairwise crossover -> mutation of a whole population
    (algorithm.py:87-120) as a composition of the batched operators above."""
    P = pop.shape[0]
    parents = pop[tournament_indices(fitness, P, k=tour_k, generator=generator)]
    parents = parents[torch.randperm(P, device=pop.device, generator=generator)]
    offspring = crossover_population(parents[..., :9].contiguous(), cxpb, generator=generator)
    return mutate_population(offspring, gen, total_gens, schedule, mut_sigma_max, mut_sigma_min,
                             mutpb, H, W, min_scale_splats, max_scale_splats, generator=generator)
