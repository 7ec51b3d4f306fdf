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


def run_search(path, target, gens_ga, iters_sa):
    """genetic_approx / simulated_annealing of the checkout at `path`, at the reference's default
    sizes (modules/config.py: 512 splats, population 32, 8 elites, 8 SA tries; work size 256x256
    here).  Returns seconds per generation / iteration beyond a 0-generation run (set-up, mask,
    initial population)."""
    import random
    use(path)
    C = importlib.import_module("modules.config")
    ga = importlib.import_module("modules.algorithm").genetic_approx
    sa = importlib.import_module("modules.annealing").simulated_annealing
    H = W = 256
    common = dict(mut_sigma_max=C.MUT_SIGMA_MAX, mut_sigma_min=C.MUT_SIGMA_MIN,
                  min_scale_splats=C.MIN_SCALE_SPLATS, max_scale_splats=C.MAX_SCALE_SPLATS,
                  k_sigma=C.K_SIGMA, mask_strength=C.MASK_STRENGTH, boost_only=C.BOOST_ONLY)

    def timed(fn, n):
        out = []
        for count in (0, 0, n):      # first call: warm-up (library load / Triton compile)
            torch.manual_seed(C.SEED); random.seed(C.SEED)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = fn(count)
            torch.cuda.synchronize(); out.append((time.perf_counter() - t0, res[1]))
        return (out[2][0] - out[1][0]) / n, out[1][1], out[2][1]

    ga_s, ga_f0, ga_f1 = timed(lambda n: ga(target, H, W, "cuda", pop_size=C.POP_SIZE, n_splats=C.N_SPLATS,
                                            generations=n, tour_k=C.TOUR_K, elite_k=C.ELITE_K, cxpb=C.CXPB,
                                            mutpb=C.MUTPB, schedule=C.SCHEDULE, **common), gens_ga)
    sa_s, sa_f0, sa_f1 = timed(lambda n: sa(target, H, W, "cuda", n_splats=C.N_SPLATS, mutpb=C.MUTPB,
                                            sigma_schedule=C.SCHEDULE, iterations=n, temp0=C.SA_T0,
                                            temp_schedule=C.SA_SCHEDULE, tries_per_iter=C.SA_TRIES_PER_ITER,
                                            **common), iters_sa)
    print(f"  {os.path.basename(path)}: GA {1.0 / ga_s:.1f} generations/s ({ga_f0:.5f} -> {ga_f1:.5f} in "
          f"{gens_ga}), SA {1.0 / sa_s:.1f} iterations/s ({sa_f0:.5f} -> {sa_f1:.5f} in {iters_sa})", flush=True)
    return {"ga_generations_per_s": 1.0 / ga_s, "ga_generations_timed": gens_ga, "ga_fitness": [ga_f0, ga_f1],
            "sa_iterations_per_s": 1.0 / sa_s, "sa_iterations_timed": iters_sa, "sa_energy": [sa_f0, sa_f1]}


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
    os.environ["TQDM_DISABLE"] = "1"
    tgt = torch.from_numpy(synth.synthetic_target_np(256, 256, 9))
    search = {"what": "genetic_approx / simulated_annealing entry points at the reference's default sizes "
                      "(256x256 work size, 512 splats, population 32, 8 elites; SA 8 tries per iteration), "
                      "marginal rate beyond a 0-generation run",
              "reference": run_search(REF, tgt, 60, 60),
              "this_library": run_search(OURS, tgt, 20000, 20000)}
    search["ga_speedup"] = search["this_library"]["ga_generations_per_s"] / search["reference"]["ga_generations_per_s"]
    search["sa_speedup"] = search["this_library"]["sa_iterations_per_s"] / search["reference"]["sa_iterations_per_s"]
    print(json.dumps(search), flush=True)
    meta = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
            "triton": importlib.import_module("triton").__version__,
            "what": "reference = unmodified josedelrey/genetic-gaussian-splats modules/ (Triton path) "
                    "run on this GPU; both sides timed through fitness_population(list) -> List[float], "
                    "wall clock, best of a few synchronised calls after a warm-up",
            "tolerances": {"image": 1e-4, "fitness_rel": 1e-5}}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"meta": meta, "rows": rows, "search": search}, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
