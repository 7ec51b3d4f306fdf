"""Importance mask: the reference's compute_importance_mask (mask.py:29-83).

Golden vectors come from the reference itself (tests/golden/make_mask_golden.py and the `mask`
arrays of make_golden.py).  CPU tests pin the plain-torch path of modules/mask.py to them; GPU
tests hold the CUDA path (ggs_importance_mask, through the C ABI) to the goldens and, at the
sizes of the BASELINE configs, to the torch path."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, load_golden

MASK_TOL = 2e-5     # absolute, on weights in [0, 1]; float32 ops in a different summation order


def mask_cases():
    z = load_golden("mask_cases")
    for name in z["names"]:
        name = str(name)
        H, W = (int(v) for v in z[name + "_hw"])
        yield name, z[name + "_image"], H, W, ast.literal_eval(str(z[name + "_kwargs"])), z[name + "_mask"]


CASES = list(mask_cases())
IDS = [c[0] for c in CASES]


@pytest.mark.parametrize("name,image,H,W,kw,want", CASES, ids=IDS)
def test_torch_path_matches_reference_goldens(name, image, H, W, kw, want):
    from oracle.torch_ref import importance_mask_torch
    got = importance_mask_torch(torch.from_numpy(image), H, W, **kw)
    assert got.shape == (H, W) and got.dtype == torch.float32
    assert np.abs(got.numpy() - want).max() <= 1e-6


def test_mask_symbols_and_workspace_size():
    from ggs_b200 import native
    lib = native.lib()
    assert lib.ggs_mask_workspace_bytes(256, 256) >= 4 * 256 * 256 * 4
    assert lib.ggs_mask_workspace_bytes(0, 5) == 0


# ------------------------------------------------------------------------------------ GPU

needs_gpu = pytest.mark.gpu


@needs_gpu
@pytest.mark.parametrize("name,image,H,W,kw,want", CASES, ids=IDS)
def test_cuda_mask_matches_reference_goldens(name, image, H, W, kw, want):
    import ggs_b200
    got = ggs_b200.importance_mask(torch.from_numpy(image).cuda(), H, W, **kw)
    assert got.is_cuda and got.shape == (H, W) and got.dtype == torch.float32
    err = np.abs(got.cpu().numpy() - want).max()
    assert err <= MASK_TOL, (name, err)


@needs_gpu
def test_cuda_mask_matches_goldens_of_the_render_cases(golden):
    # algorithm.py:42-49: the mask every GA / SA run builds for its target
    from modules.mask import compute_importance_mask
    H, W = int(golden["H"]), int(golden["W"])
    got = compute_importance_mask(torch.from_numpy(golden["target"]).cuda(), H, W,
                                  edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3, gamma=0.7,
                                  floor=0.15, smooth=3, strength=0.7)
    assert got.is_cuda
    assert np.abs(got.cpu().numpy() - golden["mask"]).max() <= MASK_TOL


@needs_gpu
@pytest.mark.parametrize("H0,W0,H,W", [(256, 256, 256, 256), (700, 500, 512, 366), (300, 1024, 75, 256),
                                        (1024, 1024, 1024, 1024)])
def test_cuda_mask_matches_torch_path_at_config_sizes(H0, W0, H, W):
    from ggs_b200 import synth
    from modules.mask import compute_importance_mask
    from oracle.torch_ref import importance_mask_torch
    image = torch.from_numpy(synth.synthetic_target_np(H0, W0, 3))
    kw = dict(edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3, gamma=0.7, floor=0.15, smooth=3,
              strength=0.7)
    want = importance_mask_torch(image, H, W, **kw)
    got = compute_importance_mask(image.cuda(), H, W, **kw)
    assert np.abs(got.cpu().numpy() - want.numpy()).max() <= MASK_TOL
    # a host tensor goes through the same kernels and comes back on the host (no CPU path)
    back = compute_importance_mask(image, H, W, **kw)
    assert not back.is_cuda and torch.equal(back, got.cpu())
    # and the masked fitness that results is the same to the fitness tolerance
    import ggs_b200
    g = torch.from_numpy(synth.new_population_np(4, 60, H, W, seed=8)).cuda()
    t = torch.from_numpy(synth.synthetic_target_np(H, W, 4)).cuda()
    f_a = ggs_b200.fitness(g, t, H, W, 3.0, weight_mask=got)
    f_b = ggs_b200.fitness(g, t, H, W, 3.0, weight_mask=want.cuda())
    np.testing.assert_allclose(f_a.cpu().numpy(), f_b.cpu().numpy(), rtol=1e-5)


@needs_gpu
def test_cuda_mask_degenerate_inputs():
    import ggs_b200
    # a constant image: every cue is flat, quantiles coincide, (t - ql) / 1e-12 clamps to 0
    flat = torch.full((40, 56, 3), 0.25).cuda()
    from oracle.torch_ref import importance_mask_torch
    kw = dict(edge_scales=(1, 2, 4), smooth=3, strength=0.7)
    want = importance_mask_torch(flat.cpu(), 40, 56, **kw)
    got = ggs_b200.importance_mask(flat, 40, 56, **kw)
    assert np.abs(got.cpu().numpy() - want.numpy()).max() <= MASK_TOL
    # 1x1 work size, scale 1 only
    one = ggs_b200.importance_mask(torch.rand(5, 7, 3).cuda(), 1, 1, edge_scales=(1,))
    assert one.shape == (1, 1) and torch.isfinite(one).all()
    # an edge scale that does not fit (avg_pool2d would fail in the reference), an even box
    with pytest.raises(ggs_b200.GgsError):
        ggs_b200.importance_mask(torch.rand(8, 8, 3).cuda(), 3, 3, edge_scales=(1, 4))
    with pytest.raises(ggs_b200.GgsError):
        ggs_b200.importance_mask(torch.rand(8, 8, 3).cuda(), 8, 8, smooth=2)


@needs_gpu
def test_cuda_mask_quantiles_are_exact_order_statistics():
    """The radix select must return the same order statistics as a sort, including ties,
    negative values and denormals (checked through the normalisation of a known plane)."""
    import ggs_b200
    # Feed a plane through the public entry with parameters that reduce the pipeline to
    # norm01(norm01-mix): compare against the torch path on awkward value distributions.
    from oracle.torch_ref import importance_mask_torch
    rng = np.random.default_rng(5)
    for trial in range(6):
        H, W = int(rng.integers(3, 90)), int(rng.integers(3, 90))
        img = rng.choice([0.0, 0.25, 0.5, 1.0], size=(H, W, 3)).astype(np.float32)   # heavy ties
        if trial % 2:
            img += rng.normal(0, 1e-3, size=img.shape).astype(np.float32)
            img = np.clip(img, 0, 1)
        kw = dict(edge_scales=(1, 2), w_edge=0.6, w_var=0.4, gamma=0.9, floor=0.1, smooth=0,
                  strength=1.0)
        want = importance_mask_torch(torch.from_numpy(img), H, W, **kw)
        got = ggs_b200.importance_mask(torch.from_numpy(img).cuda(), H, W, **kw)
        assert np.abs(got.cpu().numpy() - want.numpy()).max() <= MASK_TOL, (trial, H, W)
