"""Seeded synthetic inputs of the shapes BASELINE.json names (no network, no image files).

Genomes follow the distribution of the reference's new_population (population.py:20-46):
x,y ~ U[0,1]; sigma = s_lo + u*(s_hi - s_lo) px with u ~ Beta(0.4*8, 0.6*8) for the a-axis and
Beta(0.6*8, 0.4*8) for the b-axis, stored as log; theta ~ U[-pi,pi); rgb ~ U[0,256) and
alpha ~ U[180,256), both clamped to 255.  Generated with numpy so the CPU baseline, the tests
and the GPU arm see the same bits on any machine.
"""
from __future__ import annotations

import math

import numpy as np

MIN_SCALE_SPLATS = 3.0   # config.py:23
MAX_SCALE_SPLATS = 0.1   # config.py:24


def new_population_np(B: int, N: int, H: int, W: int, seed: int = 42,
                      min_scale: float = MIN_SCALE_SPLATS,
                      max_scale: float = MAX_SCALE_SPLATS) -> np.ndarray:
    rng = np.random.default_rng(seed)
    s_lo, s_hi = float(min_scale), float(max_scale) * float(max(H, W))
    conc, eps = 8.0, 1e-6

    def log_sigma(m):
        u = rng.beta(m * conc + eps, (1.0 - m) * conc + eps, size=(B, N, 1))
        return np.log(s_lo + u * (s_hi - s_lo))

    g = np.concatenate([
        rng.uniform(0.0, 1.0, size=(B, N, 2)),
        log_sigma(0.4), log_sigma(0.6),
        rng.uniform(-math.pi, math.pi, size=(B, N, 1)),
        rng.uniform(0.0, 256.0, size=(B, N, 3)),
        rng.uniform(180.0, 256.0, size=(B, N, 1)),
    ], axis=-1).astype(np.float32)
    g[..., 0:2] = np.clip(g[..., 0:2], 0.0, 1.0)
    g[..., 5:9] = np.clip(g[..., 5:9], 0.0, 255.0)
    return np.ascontiguousarray(g)


def late_population_np(B: int, N: int, H: int, W: int, seed: int = 42, shrink: float = 0.5,
                       alpha_lo: float = 40.0) -> np.ndarray:
    """A 'late-run' population: smaller splats and a wider alpha range than the init
    distribution, so culling statistics are not only those of generation 0."""
    g = new_population_np(B, N, H, W, seed)
    rng = np.random.default_rng(seed + 1)
    lo = math.log(MIN_SCALE_SPLATS)
    g[..., 2:4] = np.maximum(g[..., 2:4] + math.log(shrink), lo)
    g[..., 8] = rng.uniform(alpha_lo, 256.0, size=(B, N)).clip(0, 255).astype(np.float32)
    return g


def synthetic_target_np(H: int, W: int, seed: int = 0) -> np.ndarray:
    """Colour ramps + a box (edges for the mask) + seeded noise, float32 [H,W,3] in [0,1]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    base = np.stack([xx, yy, 0.5 + 0.5 * np.sin(6.0 * (xx + yy))], axis=-1)
    box = ((xx > 0.3) & (xx < 0.7) & (yy > 0.25) & (yy < 0.6)).astype(np.float64)[..., None]
    t = 0.6 * base + 0.3 * box + 0.1 * rng.uniform(size=(H, W, 3))
    return np.ascontiguousarray(np.clip(t, 0.0, 1.0).astype(np.float32))


def importance_mask_np(target: np.ndarray, strength: float = 0.7) -> np.ndarray:
    """The mask the GA loop builds (algorithm.py:42-49) for a synthetic target, computed by the
    library on the current CUDA device (ggs_importance_mask; there is no CPU path -- CPU-only
    tests use the torch restatement in oracle/torch_ref.py)."""
    import torch
    from .evaluator import importance_mask
    H, W = target.shape[:2]
    m = importance_mask(torch.from_numpy(np.ascontiguousarray(target, dtype=np.float32)).cuda(), H, W,
                        edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3, gamma=0.7, floor=0.15,
                        smooth=3, strength=strength)
    return np.ascontiguousarray(m.cpu().numpy().astype(np.float32))


def count_pairs(x0, x1, y0, y1) -> int:
    """In-AABB (pixel, splat) pairs: the algorithmic work unit."""
    x0, x1, y0, y1 = (np.asarray(a, dtype=np.int64) for a in (x0, x1, y0, y1))
    return int((np.maximum(x1 - x0 + 1, 0) * np.maximum(y1 - y0 + 1, 0)).sum())
