#!/usr/bin/env python
"""Golden fixtures from the reference's REAL GPU path (tests/golden/gpu_*.npz).

make_golden.py drives the reference on CPU tensors (Triton interpreter) because the build
container has no GPU.  This script runs on the B200 box instead, through the reference's public,
unmodified entry points on device="cuda":

  modules.encode.genome_to_renderer_batched     (encode.py:63-79)
  modules.render._preprocess_genome             (render.py:9-47)     the AABBs / conics it renders with
  modules.render.render_splats_rgb_triton       (render.py:204-252)  compiled Triton kernel
  modules.fitness.fitness_many                  (fitness.py:8-31)    three modes
  modules.mask.compute_importance_mask          (mask.py:29-83)
  modules.population.new_population, modules.genetic.mutate_individual

It needs baseline/_ref (unmodified copy of the reference, baseline/make_ref_copy.sh; git-ignored,
travels with gpurun, deleted afterwards).  Same keys as make_golden.py, so the same oracle and
CUDA parity tests replay them.  Only inputs and outputs are saved.

    gpurun -- python tests/golden/make_gpu_golden.py      # writes gpurun_out/golden_gpu/*.npz
    cp gpurun_out/golden_gpu/*.npz tests/golden/
"""
import math
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GGS_REFERENCE", os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, REF)

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

assert os.path.abspath(rrender.__file__).startswith(os.path.abspath(REF)), rrender.__file__
OUT = os.path.join(ROOT, "gpurun_out", "golden_gpu")
DEV = torch.device("cuda")
KEYS_F = ("cx", "cy", "sxx", "sxy", "syy", "rc", "gc", "bc", "a")
KEYS_I = ("x0", "x1", "y0", "y1")


def synth_target(H, W, seed):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    base = torch.stack([xx, yy, 0.5 + 0.5 * torch.sin(6.0 * (xx + yy))], dim=-1)
    box = ((xx > 0.3) & (xx < 0.7) & (yy > 0.25) & (yy < 0.6)).float().unsqueeze(-1)
    t = 0.6 * base + 0.3 * box + 0.1 * torch.rand((H, W, 3), generator=g)
    return t.clamp(0, 1).to(torch.float32).contiguous()


@torch.no_grad()
def make_case(name, axes, H, W, tile, k_sigma=3.0, seed_t=0, n_images=None):
    axes = axes.to(torch.float32).contiguous()
    B, N, C = axes.shape
    n_images = B if n_images is None else n_images
    target = synth_target(H, W, seed_t).to(DEV)
    mask = rmask.compute_importance_mask(target, H, W, edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3,
                                         gamma=0.7, floor=0.15, smooth=3,
                                         strength=rcfg.MASK_STRENGTH).contiguous()
    g = axes.to(DEV)
    chol = renc.genome_to_renderer_batched(g.clone())
    parts = [rrender._preprocess_genome(chol[b], H, W, k_sigma, DEV) for b in range(B)]
    dec = {k: torch.stack([p[k] for p in parts]).cpu().numpy() for k in KEYS_F + KEYS_I}
    imgs = rrender.render_splats_rgb_triton(chol[:n_images], H, W, k_sigma=k_sigma, device="cuda", tile=tile)
    pop = [g[b] for b in range(B)]
    f_plain = rfit.fitness_many(pop, target, H, W, k_sigma, "cuda", tile=tile)
    f_mask = rfit.fitness_many(pop, target, H, W, k_sigma, "cuda", tile=tile, weight_mask=mask)
    f_boost = rfit.fitness_many(pop, target, H, W, k_sigma, "cuda", tile=tile, weight_mask=mask,
                                boost_only=True)
    out = dict(axes=axes.numpy(), chol=chol.cpu().numpy(), target=target.cpu().numpy(),
               mask=mask.cpu().numpy(), images=imgs.cpu().numpy(), fit_plain=f_plain.cpu().numpy(),
               fit_mask=f_mask.cpu().numpy(), fit_boost=f_boost.cpu().numpy(), H=np.int32(H),
               W=np.int32(W), tile=np.int32(tile), k_sigma=np.float32(k_sigma),
               versions=np.array([f"torch {torch.__version__}", f"triton {triton.__version__}",
                                  f"numpy {np.__version__}", torch.cuda.get_device_name(0)]))
    for k in KEYS_F + KEYS_I:
        out["dec_" + k] = dec[k]
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    pairs = int(((dec["x1"] - dec["x0"] + 1) * (dec["y1"] - dec["y0"] + 1)).sum())
    print(f"{name}: B={B} N={N} {H}x{W} tile={tile} images={n_images} pairs={pairs} "
          f"fit_mask={f_mask.cpu().numpy()[:4]} -> {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


def population(B, N, H, W, seed):
    torch.manual_seed(seed)
    return rpop.new_population(B, N, H, W, rcfg.MIN_SCALE_SPLATS, rcfg.MAX_SCALE_SPLATS, device="cpu")


def main():
    random.seed(42)
    torch.manual_seed(42)
    # BASELINE config 1 shape, the reference's tile 32
    make_case("gpu_c1_128x128_n100", population(6, 100, 128, 128, 42), 128, 128, tile=32, n_images=3)
    # ragged shape, k_sigma 2.5, tile 16
    make_case("gpu_ragged_83x120_n70", population(5, 70, 83, 120, 9), 83, 120, tile=16, k_sigma=2.5,
              seed_t=1, n_images=3)
    # BASELINE config 2 shape (256x256, 500 splats, 8 neighbours), one image kept
    make_case("gpu_c2_256x256_n500", population(8, 500, 256, 256, 3), 256, 256, tile=32, seed_t=2,
              n_images=1)
    # late-run genomes: the reference's own mutation at gen == total, a few rounds
    late = population(4, 80, 96, 96, 5)
    for b in range(late.shape[0]):
        for _ in range(6):
            rgen.mutate_individual(late[b], is_elite=False, gen=100, total_gens=100,
                                   schedule=rcfg.SCHEDULE, mut_sigma_max=rcfg.MUT_SIGMA_MAX,
                                   mut_sigma_min=rcfg.MUT_SIGMA_MIN, mutpb=0.5, H=96, W=96,
                                   min_scale_splats=rcfg.MIN_SCALE_SPLATS,
                                   max_scale_splats=rcfg.MAX_SCALE_SPLATS)
    make_case("gpu_late_96x96_n80", late, 96, 96, tile=32, seed_t=4, n_images=2)
    # adversarial: out-of-range centres / colours, alpha 0, tiny and huge splats, 11 columns
    adv = population(2, 60, 96, 64, 11)
    adv = torch.cat([adv, torch.zeros(2, 60, 2)], dim=-1)
    adv[0, 0, 0:2] = torch.tensor([-0.25, 1.5])
    adv[0, 1, 8] = 0.0
    adv[0, 2, 8] = 300.0
    adv[0, 3, 5:8] = torch.tensor([-20.0, 400.0, 128.0])
    adv[0, 4, 2:4] = torch.tensor([math.log(0.5), math.log(0.5)])
    adv[0, 5, 2:4] = torch.tensor([math.log(200.0), math.log(150.0)])
    adv[0, 7, 0:2] = torch.tensor([0.0, 0.0])
    adv[0, 8, 0:2] = torch.tensor([1.0, 1.0])
    adv[1, 0, 2:4] = torch.tensor([-40.0, -40.0])
    adv[1, 1, 4] = 3.14159
    make_case("gpu_adversarial_96x64_n60", adv, 96, 64, tile=32, seed_t=3)


if __name__ == "__main__":
    main()
