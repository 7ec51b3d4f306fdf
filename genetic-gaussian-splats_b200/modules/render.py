"""Drop-in for the reference's modules/render.py (render entry: render.py:204-252).

The Triton kernel, the per-candidate torch preprocessing loop and the argsort binning are
replaced by two CUDA launches in libggs_b200.so (decode + fused tile rasteriser)."""
from __future__ import annotations

import torch

from ggs_b200 import LAYOUT_CHOLESKY
from ggs_b200 import render as _render

_DEV = 'cuda'
__all__ = ["render_splats_rgb_triton", "_DEV"]


@torch.no_grad()
def render_splats_rgb_triton(
    genomes: torch.Tensor,
    H: int, W: int, *,
    k_sigma: float = 3.0,
    device: torch.device | str | None = None,
    background=(1.0, 1.0, 1.0),
    tile: int = 64,
    num_warps: int = 8,
    num_stages: int = 3,
    use_fp16_canvas: bool = False
) -> torch.Tensor:
    """Cholesky-layout genomes [B,N,C>=9] or [N,C>=9] -> [B,H,W,3] float32 in [0,1].

    `tile`, `num_warps`, `num_stages` were Triton tuning knobs; the image does not depend on
    them (tile binning is pure culling) and they are accepted and ignored.  The canvas is
    always fp32; `use_fp16_canvas` (never set by any reference caller) is ignored too.
    """
    dev = device or _DEV
    dev = torch.device(dev) if not isinstance(dev, torch.device) else dev
    assert dev.type == "cuda", "This renderer requires a CUDA device."

    assert genomes.ndim in (2, 3), f"genomes must be [B,N,9] or [N,9], got {genomes.shape}"
    if genomes.ndim == 2:
        genomes = genomes.unsqueeze(0)
    B, N, C = genomes.shape
    assert C >= 9, "expected at least 9 genome cols"
    return _render(genomes, int(H), int(W), k_sigma=float(k_sigma), background=background,
                   layout=LAYOUT_CHOLESKY, device=dev)
