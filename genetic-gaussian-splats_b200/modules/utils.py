"""Schedules, genome clamping, frame / curve output (reference: modules/utils.py:11-151).
The two renderer callers here (render_axes_angle_to_img, prewarm_renderer) go through the
drop-in render entry, i.e. the CUDA library."""
import csv
import math
import os
from typing import Dict, Sequence

import numpy as np
import torch


@torch.no_grad()
def wrap_angle(theta: torch.Tensor) -> torch.Tensor:
    """Wrap to [-pi, pi)."""
    return (theta + np.pi) % (2 * np.pi) - np.pi


def _anneal_factor(gen, total, kind):
    """1 -> 0 over the run: cosine, linear, or exponential decay to 0.2."""
    g = max(0, min(gen, total))
    p = g / max(1, total)
    if kind == "cosine":
        f = 0.5 * (1.0 + math.cos(math.pi * p))
    elif kind == "exp":
        f = (0.2 ** (1.0 / max(1, total))) ** g
    else:  # "linear" and anything unknown
        f = 1.0 - p
    return max(0.0, f)


def build_mut_sigma(gen: int, total_gens: int, kind: str, mut_sigma_max: dict, mut_sigma_min: dict):
    f = _anneal_factor(gen, total_gens, kind)
    return {k: mut_sigma_min[k] + f * (hi - mut_sigma_min[k]) for k, hi in mut_sigma_max.items()}


def scale_log_bounds(H: int, W: int, min_scale_splats: float, max_scale_splats: float):
    return math.log(min_scale_splats), math.log(max_scale_splats * float(max(H, W)))


def clamp_genome(ind: torch.Tensor, H: int, W: int, min_scale_splats: float,
                 max_scale_splats: float) -> torch.Tensor:
    """In-place projection onto the legal genome box; works on [N,9] and on [P,N,9]."""
    lo, hi = scale_log_bounds(H, W, min_scale_splats, max_scale_splats)
    ind[..., 0:2].clamp_(0.0, 1.0)
    ind[..., 2:4].clamp_(lo, hi)
    ind[..., 4] = wrap_angle(ind[..., 4])
    ind[..., 5:9].clamp_(0.0, 255.0)
    return ind


@torch.no_grad()
def render_axes_angle_to_img(ind_axes_angle: torch.Tensor, Hsnap: int, Wsnap: int,
                             k_sigma: float, device) -> np.ndarray:
    """One individual -> uint8 [H,W,3]."""
    from ggs_b200 import LAYOUT_AXES_ANGLE, render
    G = ind_axes_angle.unsqueeze(0) if ind_axes_angle.ndim == 2 else ind_axes_angle
    # encode + decode + render + the (img * 255).astype(uint8) conversion in one evaluation on
    # the device; only H*W*3 bytes come back
    img8 = render(G, int(Hsnap), int(Wsnap), k_sigma=float(k_sigma), layout=LAYOUT_AXES_ANGLE,
                  device=device, as_uint8=True)[0]
    return img8.cpu().numpy()


@torch.no_grad()
def save_frame_png(gen: int, ind_axes_angle: torch.Tensor, pad: int, prefix: str,
                   video_dir: str, H: int, W: int, k_sigma: float, device,
                   save_video: bool = True):
    if not save_video:
        return
    from PIL import Image
    img8 = render_axes_angle_to_img(ind_axes_angle, H, W, k_sigma, device)
    Image.fromarray(img8).save(os.path.join(video_dir, f"{prefix}_{gen:0{pad}d}.png"))


@torch.no_grad()
def prewarm_renderer(H: int, W: int, k_sigma: float, device):
    """Creates the CUDA context and loads the library before the timed loop (the reference
    used this call to trigger the Triton JIT)."""
    from modules.render import render_splats_rgb_triton
    dummy = torch.tensor([[[0.5, 0.5, math.log(2.0), math.log(2.0), 0.0, 128.0, 128.0, 128.0,
                            255.0]]], device=device, dtype=torch.float32)
    for _ in range(2):
        render_splats_rgb_triton(dummy, min(8, H), min(8, W), k_sigma=k_sigma, device=device, tile=32)
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def save_loss_curve_png(curves: Dict[str, Sequence[float]], out_path: str,
                        title: str = "GA fitness over generations", xlabel: str = "Generation",
                        ylabel: str = "MSE", log_y: bool = False, dpi: int = 144) -> None:
    if not out_path:
        return
    try:
        import matplotlib.pyplot as plt
    except Exception as e:  # matplotlib is optional
        print(f"[warn] matplotlib not available, cannot save plot: {e}")
        return
    series = {k: list(v) for k, v in curves.items() if len(v) > 0}
    if not series:
        print("[warn] No values to plot")
        return
    lengths = {len(v) for v in series.values()}
    if len(lengths) != 1:
        raise ValueError(f"curves have different lengths: { {k: len(v) for k, v in series.items()} }")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    fig, ax = plt.subplots()
    for name, values in series.items():
        ax.plot(range(len(values)), values, label=name)
    ax.set(title=title, xlabel=xlabel, ylabel=ylabel)
    if log_y:
        ax.set_yscale("log")
    ax.grid(True, which="both", alpha=0.3)
    ax.legend()
    fig.tight_layout()
    fig.savefig(out_path, dpi=dpi)
    plt.close(fig)


def save_curves_csv(curves: Dict[str, Sequence[float]], out_csv_path: str) -> None:
    """CSV with header gen,<key1>,<key2>,..."""
    if not out_csv_path:
        return
    keys = list(curves.keys())
    lengths = [len(v) for v in curves.values() if len(v) > 0]
    if not lengths:
        print("[warn] No values to save to CSV")
        return
    os.makedirs(os.path.dirname(out_csv_path), exist_ok=True)
    with open(out_csv_path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["gen"] + keys)
        for i in range(lengths[0]):
            w.writerow([i] + [curves[k][i] if i < len(curves[k]) else "" for k in keys])
