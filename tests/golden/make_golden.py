#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference, torch, triton).  The
reference's public render entry asserts a CUDA device (modules/render.py:217), so
this script drives the reference's own pieces on CPU tensors, exactly as its entry
does at modules/render.py:226-252:

  modules.encode.genome_to_renderer_batched      (encode.py:63-79)   torch CPU
  modules.render._preprocess_genome              (render.py:9-47)    torch CPU
  modules.render._gpu_bin_splats_to_tiles        (render.py:51-118)  torch CPU
  modules.render._render_tile_over_kernel        (render.py:121-200) Triton, TRITON_INTERPRET=1
  modules.fitness.fitness_many                   (fitness.py:8-31)   torch CPU, with the render
                                                                      entry swapped for the CPU driver
  modules.mask.compute_importance_mask           (mask.py:29-83)     torch CPU
  modules.population.new_population              (population.py:20-46)
  modules.genetic.mutate_individual              (genetic.py:32-92)  ("late-run" genomes)

Nothing of the reference is copied into the repo: only inputs and outputs are saved.

Usage:  python tests/golden/make_golden.py            (writes tests/golden/*.npz)
"""
import os
import sys

os.environ["TRITON_INTERPRET"] = "1"  # must precede `import triton`
REF = os.environ.get("GGS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import math  # noqa: E402
import random  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402
import triton  # noqa: E402

import modules.config as rcfg  # noqa: E402
import modules.encode as renc  # noqa: E402
import modules.fitness as rfit  # noqa: E402
import modules.genetic as rgen  # noqa: E402
import modules.mask as rmask  # noqa: E402
import modules.population as rpop  # noqa: E402
import modules.render as rrender  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CPU = torch.device("cpu")
KEYS_F = ("cx", "cy", "sxx", "sxy", "syy", "rc", "gc", "bc", "a")
KEYS_I = ("x0", "x1", "y0", "y1")


@torch.no_grad()
def ref_render_cpu(genomes, H, W, *, k_sigma=3.0, device=None, background=(1.0, 1.0, 1.0),
                   tile=64, num_warps=8, num_stages=3, use_fp16_canvas=False, _keep=None):
    """The body of render_splats_rgb_triton (render.py:219-252) on CPU tensors."""
    assert genomes.ndim in (2, 3)
    if genomes.ndim == 2:
        genomes = genomes.unsqueeze(0)
    B, N, C = genomes.shape
    assert C >= 9
    parts = [rrender._preprocess_genome(genomes[b], H, W, k_sigma, CPU) for b in range(B)]
    cat = {k: torch.cat([p[k] for p in parts], dim=0) for k in parts[0].keys()}
    flat_idx, tile_off, tile_cnt, nTX, nTY, ntiles = rrender._gpu_bin_splats_to_tiles(
        cat["x0"], cat["x1"], cat["y0"], cat["y1"], B, N, H, W, tile)
    img = torch.empty((B, H, W, 3), dtype=torch.float32).contiguous()
    img[:] = torch.as_tensor(background, dtype=torch.float32)
    sb, sh, sw, _ = img.stride()
    rrender._render_tile_over_kernel[(B * ntiles,)](
        img, H, W, sb, sh, sw,
        cat["cx"], cat["cy"], cat["sxx"], cat["sxy"], cat["syy"],
        cat["rc"], cat["gc"], cat["bc"], cat["a"],
        cat["x0"], cat["x1"], cat["y0"], cat["y1"],
        flat_idx, tile_off, tile_cnt, nTX, ntiles,
        TILE_W=tile, TILE_H=tile, num_warps=num_warps, num_stages=num_stages)
    if _keep is not None:
        _keep.update({k: v.reshape(B, N).numpy().copy() for k, v in cat.items()})
    return img.clamp_(0.0, 1.0).to(torch.float32)


def synth_target(H, W, seed):
    """Smooth colour ramps plus seeded noise: has edges for the mask, values in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    base = torch.stack([xx, yy, 0.5 + 0.5 * torch.sin(6.0 * (xx + yy))], dim=-1)
    box = ((xx > 0.3) & (xx < 0.7) & (yy > 0.25) & (yy < 0.6)).float().unsqueeze(-1)
    t = 0.6 * base + 0.3 * box + 0.1 * torch.rand((H, W, 3), generator=g)
    return t.clamp(0, 1).to(torch.float32).contiguous()


def ref_mask(target, H, W):
    # the call made at algorithm.py:42-49, mask_strength = config.MASK_STRENGTH
    return rmask.compute_importance_mask(target, H, W, edge_scales=(1, 2, 4), w_edge=0.7,
                                         w_var=0.3, gamma=0.7, floor=0.15, smooth=3,
                                         strength=rcfg.MASK_STRENGTH).contiguous()


def make_case(name, axes, H, W, tile, k_sigma=3.0, seed_t=0, extra=None):
    axes = axes.to(torch.float32).contiguous()
    B, N, C = axes.shape
    target = synth_target(H, W, seed_t)
    mask = ref_mask(target, H, W)
    chol = renc.genome_to_renderer_batched(axes.clone())
    keep = {}
    imgs = ref_render_cpu(chol, H, W, k_sigma=k_sigma, tile=tile, _keep=keep)

    saved = rfit.render_splats_rgb_triton
    rfit.render_splats_rgb_triton = ref_render_cpu
    try:
        pop = [axes[b] for b in range(B)]
        f_plain = rfit.fitness_many(pop, target, H, W, k_sigma, CPU, tile=tile)
        f_mask = rfit.fitness_many(pop, target, H, W, k_sigma, CPU, tile=tile, weight_mask=mask)
        f_boost = rfit.fitness_many(pop, target, H, W, k_sigma, CPU, tile=tile, weight_mask=mask,
                                    boost_only=True)
    finally:
        rfit.render_splats_rgb_triton = saved

    out = dict(axes=axes.numpy(), chol=chol.numpy(), target=target.numpy(), mask=mask.numpy(),
               images=imgs.numpy(), fit_plain=f_plain.numpy(), fit_mask=f_mask.numpy(),
               fit_boost=f_boost.numpy(), H=np.int32(H), W=np.int32(W), tile=np.int32(tile),
               k_sigma=np.float32(k_sigma),
               versions=np.array([f"torch {torch.__version__}", f"triton {triton.__version__}",
                                  f"numpy {np.__version__}"]))
    for k in KEYS_F + KEYS_I:
        out["dec_" + k] = keep[k]
    if extra:
        out.update(extra)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    pairs = int(((keep["x1"] - keep["x0"] + 1) * (keep["y1"] - keep["y0"] + 1)).sum())
    print(f"{name}: B={B} N={N} {H}x{W} tile={tile} pairs={pairs} "
          f"fit_plain={f_plain.numpy()} -> {os.path.getsize(path)/1024:.0f} KiB", flush=True)


def population(B, N, H, W, seed):
    torch.manual_seed(seed)
    return rpop.new_population(B, N, H, W, rcfg.MIN_SCALE_SPLATS, rcfg.MAX_SCALE_SPLATS,
                               device="cpu")


def main():
    random.seed(42)
    torch.manual_seed(42)

    # 1. BASELINE config 1 shape (128x128, 100 splats), reference tile 32, seed 42.
    make_case("c1_128x128_n100", population(4, 100, 128, 128, 42), 128, 128, tile=32)

    # 2. ragged: H, W not multiples of any tile size, H != W, tile 16.
    make_case("ragged_50x44_n30", population(3, 30, 50, 44, 7), 50, 44, tile=16, seed_t=1)

    # 3. the prewarm call (utils.py:73-82): 8x8, one splat, tile 32 > image.  The prewarm
    #    genome is Cholesky-layout; expressed here in axes-angle with theta=0 so it encodes
    #    to itself up to exp/log rounding.
    dummy = torch.tensor([[[0.5, 0.5, math.log(2.0), math.log(2.0), 0.0, 128.0, 128.0, 128.0,
                            255.0]]], dtype=torch.float32)
    make_case("prewarm_8x8_n1", dummy, 8, 8, tile=32, seed_t=2)

    # 4. adversarial genomes: out-of-range centres/colours, alpha 0, needle and huge splats,
    #    extra genome columns (C = 11), k_sigma = 2.5.
    adv = population(2, 60, 96, 64, 11)
    adv = torch.cat([adv, torch.zeros(2, 60, 2)], dim=-1)  # extra columns are ignored
    adv[0, 0, 0:2] = torch.tensor([-0.25, 1.5])            # centre outside the image -> clamped
    adv[0, 1, 8] = 0.0                                      # alpha 0
    adv[0, 2, 8] = 300.0                                    # alpha > 255 -> clamped
    adv[0, 3, 5:8] = torch.tensor([-20.0, 400.0, 128.0])    # colours out of range
    adv[0, 4, 2:4] = torch.tensor([math.log(0.05), math.log(0.05)])   # sub-pixel splat
    adv[0, 5, 2:4] = torch.tensor([math.log(200.0), math.log(150.0)])  # covers everything
    adv[0, 6, 2:4] = torch.tensor([math.log(40.0), math.log(0.3)])     # needle
    adv[0, 6, 4] = 0.7
    adv[0, 7, 0:2] = torch.tensor([0.0, 0.0])               # corner
    adv[0, 8, 0:2] = torch.tensor([1.0, 1.0])               # opposite corner
    adv[1, 0, 2:4] = torch.tensor([-40.0, -40.0])           # exp underflow -> 1e-6 floors
    adv[1, 1, 4] = 3.14159                                  # theta at the wrap
    make_case("adversarial_96x64_n60", adv, 96, 64, tile=32, k_sigma=2.5, seed_t=3)

    # 5. late-run genomes: init population pushed through the reference's own mutation at
    #    gen == total (smallest sigmas), a few rounds, so culling statistics differ from init.
    late = population(2, 50, 64, 64, 5)
    for b in range(late.shape[0]):
        for _ in range(6):
            rgen.mutate_individual(late[b], is_elite=False, gen=100, total_gens=100,
                                   schedule=rcfg.SCHEDULE, mut_sigma_max=rcfg.MUT_SIGMA_MAX,
                                   mut_sigma_min=rcfg.MUT_SIGMA_MIN, mutpb=0.5, H=64, W=64,
                                   min_scale_splats=rcfg.MIN_SCALE_SPLATS,
                                   max_scale_splats=rcfg.MAX_SCALE_SPLATS)
    make_case("late_64x64_n50", late, 64, 64, tile=32, seed_t=4)

    # 6. tile invariance: case 2's genomes rendered with the reference at tile 32 as well.
    g2 = population(3, 30, 50, 44, 7)
    chol = renc.genome_to_renderer_batched(g2.clone())
    img32 = ref_render_cpu(chol, 50, 44, tile=32)
    img16 = ref_render_cpu(chol, 50, 44, tile=16)
    print("tile 16 vs 32 max abs diff:", float((img32 - img16).abs().max()))


if __name__ == "__main__":
    main()
