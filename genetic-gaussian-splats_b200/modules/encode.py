"""Drop-in for the reference's modules/encode.py (encode.py:5-24, :28-59, :63-79).

The axes-angle -> Cholesky conversion runs as one CUDA kernel (ggs_encode) with the
reference's fp32 operation order.  Inside fitness_many it is fused into the decode and
this module is not on the path; it serves the callers that render single images."""
from __future__ import annotations

import torch

from ggs_b200 import encode as _encode


def _on_cuda(t: torch.Tensor):
    return t if t.is_cuda else t.to("cuda")


@torch.no_grad()
def axes_angle_to_cholesky(a_log: torch.Tensor, b_log: torch.Tensor, theta: torch.Tensor):
    """(log sigma_x, log sigma_y, theta) -> (log l11, log l22, l21), elementwise."""
    shape, src = a_log.shape, a_log.device
    rows = torch.zeros((a_log.numel(), 9), dtype=torch.float32, device=_on_cuda(a_log).device)
    rows[:, 2] = a_log.reshape(-1)
    rows[:, 3] = b_log.reshape(-1)
    rows[:, 4] = theta.reshape(-1)
    out = _encode(rows)
    return (out[:, 2].reshape(shape).to(src), out[:, 3].reshape(shape).to(src),
            out[:, 4].reshape(shape).to(src))


@torch.no_grad()
def genome_to_renderer(ind_axes_angle: torch.Tensor) -> torch.Tensor:
    """[N,C>=9] (or [C]) axes-angle genome -> [N,9] Cholesky layout, colours clamped."""
    if ind_axes_angle.ndim == 1:
        ind_axes_angle = ind_axes_angle.unsqueeze(0)
    src = ind_axes_angle.device
    return _encode(_on_cuda(ind_axes_angle)).to(device=src, dtype=ind_axes_angle.dtype)


@torch.no_grad()
def genome_to_renderer_batched(G_axes: torch.Tensor) -> torch.Tensor:
    """[B,N,C>=9] -> [B,N,9]."""
    B, N, C = G_axes.shape
    src = G_axes.device
    return _encode(_on_cuda(G_axes)).to(device=src, dtype=G_axes.dtype).reshape(B, N, 9)
