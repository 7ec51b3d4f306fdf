#!/usr/bin/env python
"""The reference's REAL GPU path against this library, on the same B200 and the same inputs.

Needs baseline/_ref (an unmodified copy of the reference made by baseline/make_ref_copy.sh; see
baseline/README.md) and a GPU.  For every shape: fitness_many of the reference (encode ->
render_splats_rgb_triton -> squared error) and of modules/ here, the three fitness modes, the
rendered images, the importance mask; reports max |image diff|, max relative fitness diff,
whether the rankings agree, and the time per call of both (wall clock around a synchronised
call: the reference's call contains host syncs, so CUDA events would undercount it).

    python tools/reference_gpu_compare.py [--out gpurun_out/reference_gpu_compare.json]
"""
import argparse, importlib, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
OURS = os.path.join(ROOT, "genetic-gaussian-splats_b200")
sys.path.insert(0, ROOT)

import numpy as np
import torch

SHAPES = [  # name, side, splats, candidates, images compared, reference timing repeats
    ("config 1: 128x128, 100 splats, population 32", 128, 100, 32, 32, 3),
    ("config 2: 256x256, 500 splats, 8 neighbours", 256, 500, 8, 8, 3),
    ("config 3: 256x256, 1,000 splats, population 1,024", 256, 1000, 1024, 16, 2),
    ("512x512, 4,000 splats, 64 candidates (config 4 shape)", 512, 4000, 64, 4, 2),
]


def use(path):
    """Make `modules` resolve to the checkout at `path` (both trees use that package name)."""
    for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
        del sys.modules[k]
    for p in (REF, OURS):
        while p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, path)
    importlib.invalidate_caches()
    fit = importlib.import_module("modules.fitness")
    assert os.path.dirname(os.path.dirname(os.path.abspath(fit.__file__))) == path, fit.__file__
    return (fit, importlib.import_module("modules.render"), importlib.import_module("modules.encode"),
            importlib.import_module("modules.mask"))


def wall(fn, repeats):
    fn(); torch.cuda.synchronize()   # warm-up (Triton compile on the reference side)
    best = float("inf")
    for _ in range(repeats):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def run_side(path, inputs, repeats_scale):
    fit, render, encode, mask = use(path)
    out = []
    for (name, side, N, B, n_img, reps), (g, t, src) in zip(SHAPES, inputs):
        H = W = side
        pop = list(torch.from_numpy(g).cuda().unbind(0))
        target = torch.from_numpy(t).cuda()
        m = mask.compute_importance_mask(torch.from_numpy(src).cuda(), H, W, edge_scales=(1, 2, 4), w_edge=0.7,
                                         w_var=0.3, gamma=0.7, floor=0.15, smooth=3, strength=0.7)
        r = {"mask": m.float().cpu().numpy()}
        r["fit_plain"] = fit.fitness_many(pop, target, H, W, 3.0, "cuda", tile=32).cpu().numpy()
        r["fit_mask"] = fit.fitness_many(pop, target, H, W, 3.0, "cuda", tile=32, weight_mask=m).cpu().numpy()
        r["fit_boost"] = fit.fitness_many(pop, target, H, W, 3.0, "cuda", tile=32, weight_mask=m,
                                          boost_only=True).cpu().numpy()
        chol = encode.genome_to_renderer_batched(torch.stack(pop[:n_img]))
        r["images"] = render.render_splats_rgb_triton(chol, H, W, k_sigma=3.0, device="cuda", tile=32).cpu().numpy()
        r["seconds"] = wall(lambda: fit.fitness_population(pop, target, H, W, 3.0, "cuda", tile=32,
                                                           weight_mask=m), reps * repeats_scale)
        out.append(r)
        print(f"  {os.path.basename(path)}: {name}: {r['seconds'] * 1e3:.3f} ms per fitness_population call",
              flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_gpu_compare.json"))
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "modules")):
        sys.exit(f"{REF} missing: run baseline/make_ref_copy.sh in the build container first")
    assert torch.cuda.is_available()
    sys.path.insert(0, OURS)
    from ggs_b200 import synth
    sys.path.remove(OURS)
    inputs = []
    for name, side, N, B, n_img, reps in SHAPES:
        g = synth.new_population_np(B, N, side, side, seed=42)
        t = synth.synthetic_target_np(side, side, 0)
        src = synth.synthetic_target_np(side + side // 2, side + side // 3, 5)  # mask source: resized
        inputs.append((g, t, src))
    ref = run_side(REF, inputs, 1)
    ours = run_side(OURS, inputs, 5)
    rows = []
    for (name, side, N, B, n_img, reps), a_, b_ in zip(SHAPES, ref, ours):
        row = {"shape": name, "candidates": B,
               "image_max_abs_diff": float(np.abs(a_["images"] - b_["images"]).max()),
               "mask_max_abs_diff": float(np.abs(a_["mask"] - b_["mask"]).max())}
        for mode in ("plain", "mask", "boost"):
            fa, fb = a_["fit_" + mode].astype(np.float64), b_["fit_" + mode].astype(np.float64)
            row[f"fitness_{mode}_max_rel_diff"] = float(np.abs(fb / fa - 1.0).max())
            ra, rb = np.argsort(fa, kind="stable"), np.argsort(fb, kind="stable")
            row[f"ranking_{mode}_identical"] = bool(np.array_equal(ra, rb))
            for k in (8, 32):
                row[f"top{k}_{mode}_identical"] = bool(np.array_equal(ra[:k], rb[:k]))
            # where the full rankings differ: how far apart (relative) the reference's own values
            # of the swapped candidates are -- a swap inside the 1e-5 tolerance is a tie
            swapped = np.nonzero(ra != rb)[0]
            row[f"ranking_{mode}_positions_differing"] = int(swapped.size)
            row[f"ranking_{mode}_max_rel_gap_of_swapped"] = (
                float(np.abs(fa[ra[swapped]] / fa[rb[swapped]] - 1.0).max()) if swapped.size else 0.0)
        row["reference_ms_per_call"] = a_["seconds"] * 1e3
        row["this_library_ms_per_call"] = b_["seconds"] * 1e3
        row["reference_candidates_per_s"] = B / a_["seconds"]
        row["this_library_candidates_per_s"] = B / b_["seconds"]
        row["speedup"] = a_["seconds"] / b_["seconds"]
        rows.append(row)
        print(json.dumps(row), flush=True)
    meta = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
            "triton": importlib.import_module("triton").__version__,
            "what": "reference = unmodified josedelrey/genetic-gaussian-splats modules/ (Triton path) "
                    "run on this GPU; both sides timed through fitness_population(list) -> List[float], "
                    "wall clock, best of a few synchronised calls after a warm-up",
            "tolerances": {"image": 1e-4, "fitness_rel": 1e-5}}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"meta": meta, "rows": rows}, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
