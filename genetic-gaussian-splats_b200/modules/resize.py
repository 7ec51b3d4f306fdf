"""Work-size helpers (reference: modules/resize.py:6-20).  One-time host-side scalar maths."""
from typing import Tuple

import numpy as np
import torch


def choose_work_size(Ht: int, Wt: int, max_side: int = 128) -> Tuple[int, int]:
    """Scale (Ht, Wt) so the longer side equals max_side, keeping the aspect ratio."""
    if Ht >= Wt:
        return max_side, max(1, int(round(Wt * max_side / Ht)))
    return max(1, int(round(Ht * max_side / Wt))), max_side


def scale_genome_pixels_anisotropic(ind: torch.Tensor, sH: float, sW: float) -> torch.Tensor:
    """Rescale the pixel-space sigmas of an axes-angle genome for a render at another size:
    log sigma_x += log sW, log sigma_y += log sH (positions are already relative)."""
    out = ind.clone()
    out[:, 2] += float(np.log(sW))
    out[:, 3] += float(np.log(sH))
    return out
