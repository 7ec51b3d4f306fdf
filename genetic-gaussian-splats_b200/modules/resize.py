"""Work-size helpers, host-side scalar maths run once per run (reference: modules/resize.py:6-20)."""
import math
from typing import Tuple

import torch


def choose_work_size(Ht: int, Wt: int, max_side: int = 128) -> Tuple[int, int]:
    """(H, W) with the longer side equal to max_side and the aspect ratio of (Ht, Wt)."""
    long_side, short_side = max(Ht, Wt), min(Ht, Wt)
    short = max(1, int(round(short_side * max_side / long_side)))
    return (max_side, short) if Ht >= Wt else (short, max_side)


def scale_genome_pixels_anisotropic(ind: torch.Tensor, sH: float, sW: float) -> torch.Tensor:
    """Copy of an axes-angle genome whose pixel-space sigmas are rescaled for a render at
    another size: log sigma_x gains log sW, log sigma_y gains log sH; positions are relative
    already."""
    scaled = ind.clone()
    scaled[:, 2].add_(math.log(sW))
    scaled[:, 3].add_(math.log(sH))
    return scaled
