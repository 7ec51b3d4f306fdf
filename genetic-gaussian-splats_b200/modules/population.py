"""Population initialisation (reference: modules/population.py:6-59).

Genome row (axes-angle): x, y in [0,1]; log sigma_x, log sigma_y (pixels); theta in [-pi, pi);
r, g, b, alpha in [0,255].  The population is one resident [P,N,9] tensor; the list form the
reference passes around is a list of views into it."""
import math
from typing import List

import torch


def sample_log_scales_beta_linear(B, N, s_lo, s_hi, m=0.5, concentration=8.0, device='cuda',
                                  dtype=torch.float32):
    """log of sigma = s_lo + u (s_hi - s_lo), u ~ Beta(m c, (1-m) c): [B,N,1]."""
    eps = 1e-6
    c = max(concentration, eps)
    a = torch.tensor(m * c + eps, device=device, dtype=dtype)
    b = torch.tensor((1.0 - m) * c + eps, device=device, dtype=dtype)
    u = torch.distributions.Beta(a, b).sample((B, N, 1))
    return (s_lo + u * (s_hi - s_lo)).log()


@torch.no_grad()
def new_population(batch_size: int, n_splats: int, H: int, W: int,
                   min_scale_splats: float, max_scale_splats: float,
                   device='cuda', dtype=torch.float32) -> torch.Tensor:
    B, N = batch_size, n_splats
    s_lo = float(min_scale_splats)
    s_hi = float(max_scale_splats * float(max(H, W)))

    def uniform(cols, lo, hi):
        return torch.empty(B, N, cols, device=device, dtype=dtype).uniform_(lo, hi)

    G = torch.cat([
        uniform(2, 0.0, 1.0),
        sample_log_scales_beta_linear(B, N, s_lo, s_hi, m=0.4, concentration=8.0, device=device, dtype=dtype),
        sample_log_scales_beta_linear(B, N, s_lo, s_hi, m=0.6, concentration=8.0, device=device, dtype=dtype),
        uniform(1, -math.pi, math.pi),
        uniform(3, 0.0, 256.0),
        uniform(1, 180.0, 256.0),
    ], dim=-1)
    G[..., 0:2].clamp_(0.0, 1.0)
    G[..., 5:9].clamp_(0.0, 255.0)
    return G


def new_individual(n_splats: int, H: int, W: int, min_scale_splats: float,
                   max_scale_splats: float, device='cuda') -> torch.Tensor:
    return new_population(1, n_splats, H, W, min_scale_splats, max_scale_splats, device=device)[0]


def duplicate_individual(ind: torch.Tensor) -> torch.Tensor:
    return ind.clone()


def population_to_list(pop_tensor: torch.Tensor) -> List[torch.Tensor]:
    return list(pop_tensor.unbind(0))
