#!/usr/bin/env python
"""Golden vectors for the importance mask, made by the REFERENCE's own
modules/mask.py::compute_importance_mask (mask.py:29-83) on torch CPU in this container.

    python tests/golden/make_mask_golden.py        # writes tests/golden/mask_cases.npz

Cases cover what the goldens of make_golden.py do not: a source image larger / smaller than the
work size (the bilinear resize of mask.py:47), uint8-range input (the /255 branch, mask.py:45),
other edge scales, no smoothing, a 5x5 smoothing box, strength 1.0 and odd sizes.  The reference
is imported from /root/reference and never copied."""
import os
import sys

REF = os.environ.get("GGS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import modules.mask as rmask  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def picture(H, W, seed):
    """Ramps, a box, a disc and noise: edges at several scales, flat areas, values in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    base = torch.stack([xx, yy * yy, 0.5 + 0.5 * torch.cos(9.0 * xx - 4.0 * yy)], dim=-1)
    box = ((xx > 0.2) & (xx < 0.55) & (yy > 0.3) & (yy < 0.8)).float()[..., None]
    disc = (((xx - 0.7) ** 2 + (yy - 0.35) ** 2) < 0.03).float()[..., None]
    t = 0.5 * base + 0.25 * box + 0.2 * disc * torch.tensor([1.0, 0.2, 0.6]) \
        + 0.08 * torch.rand((H, W, 3), generator=g)
    return t.clamp(0, 1).to(torch.float32).contiguous()


CASES = [
    # name, H0, W0, H, W, scale255, kwargs
    ("down", 90, 70, 48, 40, False, dict(edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3, gamma=0.7,
                                          floor=0.15, smooth=3, strength=0.7)),
    ("up", 20, 30, 64, 61, False, dict(edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3, gamma=0.7,
                                        floor=0.15, smooth=3, strength=0.7)),
    ("u8range", 57, 83, 57, 83, True, dict(edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3,
                                            gamma=0.7, floor=0.15, smooth=3, strength=0.7)),
    ("defaults", 64, 96, 64, 96, False, dict()),       # smooth 0, strength 1.0 (mask.py:31-38)
    ("scales13", 75, 50, 75, 50, False, dict(edge_scales=(1, 3), w_edge=0.5, w_var=0.5, gamma=1.3,
                                              floor=0.0, smooth=5, strength=1.0)),
    ("tiny", 9, 11, 9, 11, False, dict(edge_scales=(1, 2, 4), smooth=3, strength=0.7)),
]


def main():
    out = {}
    for i, (name, H0, W0, H, W, u8, kw) in enumerate(CASES):
        img = picture(H0, W0, 10 + i)
        if u8:
            img = (img * 255.0).round()
        with torch.no_grad():
            m = rmask.compute_importance_mask(img, H, W, **kw).contiguous()
        assert m.shape == (H, W)
        out[f"{name}_image"] = img.numpy()
        out[f"{name}_mask"] = m.numpy()
        out[f"{name}_hw"] = np.array([H, W], dtype=np.int32)
        out[f"{name}_kwargs"] = np.array(repr(kw))
        print(name, tuple(img.shape), "->", tuple(m.shape), "min %.4f max %.4f" % (m.min(), m.max()))
    out["names"] = np.array([c[0] for c in CASES])
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "mask_cases.npz"), **out)


if __name__ == "__main__":
    main()
