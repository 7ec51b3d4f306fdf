"""Importance (edge/detail) weight mask -- counterpart of the reference's modules/mask.py
(compute_importance_mask, mask.py:29-83), the `weight_mask` input of the fitness.

Everything runs in the library (ggs_importance_mask: resize, luma, multi-scale Sobel, local
variance, exact-order-statistic quantiles, smoothing, shaping -- all on the device, no host
round trip).  There is no CPU path: a host tensor is uploaded, the mask is computed on the
current CUDA device and returned on the input's device, as the reference's callers expect
(algorithm.py:42-49 moves it with `.to(device)` afterwards).  The torch restatement that serves
as the kernels' fp32 reference lives with the tests (oracle/torch_ref.py)."""
from __future__ import annotations

import torch


@torch.no_grad()
def compute_importance_mask(
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
    from ggs_b200 import importance_mask
    assert torch.cuda.is_available(), "compute_importance_mask needs a CUDA device (no CPU path)"
    src = target_hw3 if target_hw3.is_cuda else target_hw3.cuda()
    out = importance_mask(src, H, W, edge_scales=edge_scales, w_edge=w_edge, w_var=w_var,
                          gamma=gamma, floor=floor, smooth=smooth, strength=strength)
    return out if target_hw3.is_cuda else out.to(target_hw3.device)
