"""Genetic-algorithm loop (reference: modules/algorithm.py:17-195), same entry point and
return value, restructured around a population tensor that stays resident on the device:
selection, crossover and mutation are one CUDA launch (ggs_ga_breed via modules/genetic.py),
elitism is a row copy, the evaluation is the fused CUDA path, and the only host transfer per
generation is three statistics of the fitness vector.  Launched under torchrun the same
function shards the evaluation over the ranks (BASELINE config 4): every rank holds the whole
population and breeds the same next generation from the same counter-based random stream,
evaluates its contiguous slice, and one NCCL all-gather moves the fitness vector;
rank 0 alone writes frames and curves.
Two reference quirks are dropped because they cannot change results: elites are not
re-evaluated (the evaluation is deterministic, algorithm.py:134) -- their stored fitness is
reused -- and the offspring that elitism discards are not evaluated at all."""
from typing import Tuple

import torch
import torch.nn.functional as F

try:
    from tqdm.auto import tqdm
except Exception:  # tqdm is optional
    def tqdm(it, **_):
        return it

from modules.fitness import fitness_many
from modules.genetic import breed_population
from modules.mask import compute_importance_mask
from modules.population import new_population
from ggs_b200.distributed import ShardedEvaluator, init_from_env, replicate
from modules.utils import (_anneal_factor, prewarm_renderer, save_curves_csv, save_frame_png,
                           save_loss_curve_png)


def prepare_target(target_img_uint8: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """float32 [H,W,3] in [0,1], bilinearly resized to the work size (algorithm.py:33-39)."""
    t = target_img_uint8.to(torch.float32)
    if t.max() > 1.5:
        t = t / 255.0
    if t.shape[0] != H or t.shape[1] != W:
        t = F.interpolate(t.permute(2, 0, 1).unsqueeze(0), size=(H, W), mode='bilinear',
                          align_corners=False)[0].permute(1, 2, 0)
    return t.contiguous()


@torch.no_grad()
def genetic_approx(target_img_uint8: torch.Tensor,
                   H: int, W: int, device,
                   pop_size: int, n_splats: int, generations: int,
                   tour_k: int, elite_k: int, cxpb: float, mutpb: float,
                   mut_sigma_max: dict, mut_sigma_min: dict, schedule: str,
                   min_scale_splats: float, max_scale_splats: float,
                   k_sigma: float, mask_strength: float, boost_only: bool,
                   save_video: bool = False, frame_every: int = 5000,
                   video_dir: str = "", prefix: str = "ga",
                   loss_png_path: str = "",
                   loss_csv_path: str = "",
                   loss_log_y: bool = False) -> Tuple[torch.Tensor, float]:
    rank, world = init_from_env()      # (0, 1) unless launched under torchrun
    if rank != 0:                      # one writer
        save_video, loss_png_path, loss_csv_path = False, "", ""
    target = prepare_target(target_img_uint8, H, W).to(device)
    # the mask is built on the device too (ggs_importance_mask); target and mask stay resident
    imp_mask = compute_importance_mask(target, H, W, edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3,
                                       gamma=0.7, floor=0.15, smooth=3, strength=mask_strength)
    prewarm_renderer(H, W, k_sigma, device)

    if world > 1:
        sharded = ShardedEvaluator(target, H, W, k_sigma, weight_mask=imp_mask,
                                   boost_only=boost_only, device=target.device)
        evaluate = sharded.fitness     # this rank's slice + all-gather: the full vector everywhere
    else:
        def evaluate(pop_tensor: torch.Tensor) -> torch.Tensor:
            return fitness_many(pop_tensor, target, H, W, k_sigma, device, tile=32,
                                weight_mask=imp_mask, boost_only=boost_only)

    pop = replicate(new_population(pop_size, n_splats, H, W, min_scale_splats, max_scale_splats,
                                   device=device))
    fit = evaluate(pop)

    def summarise(fit_vec: torch.Tensor):
        """(stable order, [best, mean, median] on the host): one small transfer per generation
        instead of the whole fitness vector (statistics.median semantics: mean of the middle two)."""
        order = torch.argsort(fit_vec, stable=True)
        ranked = fit_vec[order].double()
        n = ranked.shape[0]
        stats = torch.stack([ranked[0], ranked.mean(), 0.5 * (ranked[(n - 1) // 2] + ranked[n // 2])])
        return order, stats.tolist()

    order, (best_fit, mean_fit, median_fit) = summarise(fit)
    best_ind = pop[order[0]].clone()
    no_improve = 0
    curves = {"best": [best_fit], "mean": [mean_fit], "median": [median_fit]}

    pad = len(str(generations))
    if save_video:
        save_frame_png(0, best_ind, pad, prefix, video_dir, H, W, k_sigma, device, save_video)

    n_elite = max(1, elite_k)
    keep = pop_size - n_elite
    # Two generation buffers used alternately, each [n_elite + P, N, 9]: the breeding kernel writes
    # the children behind the elite rows, so the next population (elites first, then the first
    # `keep` children, algorithm.py:128-141) is a view of the buffer: no concatenation, no copy
    # of the children.
    room = [torch.empty((n_elite + pop_size, n_splats, 9), dtype=torch.float32, device=pop.device)
            for _ in range(2)]
    run_seed = int(replicate(torch.randint(0, 2**31 - 1, (1,)).to(device)).item())  # follows torch.manual_seed
    pbar = tqdm(range(1, generations + 1), desc="GA generations", leave=True,
                **({"disable": True} if rank != 0 else {}))
    try:
        for gen in pbar:
            nxt = room[gen & 1]
            # selection -> crossover -> mutation: one launch on the resident tensor
            children = breed_population(pop, fit, gen, generations, schedule, mut_sigma_max,
                                        mut_sigma_min, tour_k, cxpb, mutpb, H, W,
                                        min_scale_splats, max_scale_splats, seed=run_seed,
                                        out=nxt[n_elite:])
            off_fit = evaluate(children[:keep])

            # elitism: the n_elite best of the current generation survive unchanged
            elite_idx = order[:n_elite]
            nxt[:n_elite] = pop[elite_idx][..., :9]
            fit = torch.cat([fit[elite_idx], off_fit], dim=0)
            pop = nxt[:pop_size]

            order, (gen_best, mean_fit, median_fit) = summarise(fit)
            if gen_best + 1e-10 < best_fit:
                best_fit, best_ind, no_improve = gen_best, pop[order[0]].clone(), 0
            else:
                no_improve += 1
            curves["best"].append(best_fit)
            curves["mean"].append(mean_fit)
            curves["median"].append(median_fit)

            if save_video and gen % max(1, frame_every) == 0:
                save_frame_png(gen, best_ind, pad, prefix, video_dir, H, W, k_sigma, device, save_video)
            if hasattr(pbar, "set_postfix"):
                pbar.set_postfix(best_mse=f"{best_fit:.6f}", stale=no_improve,
                                 sigma_fac=f"{_anneal_factor(gen, generations, schedule):.3f}")
    except KeyboardInterrupt:
        print("\n[Interrupted] Returning current best individual...", flush=True)
    finally:
        if hasattr(pbar, "close"):
            pbar.close()

    try:
        save_loss_curve_png(curves, loss_png_path, title=f"{prefix} fitness", xlabel="Generation",
                            ylabel="MSE", log_y=loss_log_y, dpi=144)
        save_curves_csv(curves, loss_csv_path)
        if loss_png_path:
            print(f"Saved loss plot to {loss_png_path}")
        if loss_csv_path:
            print(f"Saved loss CSV to {loss_csv_path}")
    except Exception as e:
        print(f"[warn] Could not save loss curves: {e}")
    return best_ind.cpu(), best_fit
