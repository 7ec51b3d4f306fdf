"""Importance (edge/detail) weight mask -- counterpart of the reference's modules/mask.py
(compute_importance_mask, mask.py:29-83), the `weight_mask` input of the fitness.  A target on
a CUDA device goes through the library (ggs_importance_mask: resize, luma, multi-scale Sobel,
local variance, exact-order-statistic quantiles, smoothing, shaping -- all on the device, no
host round trip); a CPU target takes the plain torch ops below, which follow the reference
one to one and double as the fp32 reference of the CUDA path in the tests."""
from __future__ import annotations

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
    if target_hw3.is_cuda:
        from ggs_b200 import importance_mask
        return importance_mask(target_hw3, H, W, edge_scales=edge_scales, w_edge=w_edge,
                               w_var=w_var, gamma=gamma, floor=floor, smooth=smooth,
                               strength=strength)
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
