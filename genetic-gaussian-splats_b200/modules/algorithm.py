"""Genetic-algorithm loop (reference: modules/algorithm.py:17-195), same entry point and
return value, restructured around a population tensor that stays resident on the device:
selection, crossover and mutation are one CUDA launch (ggs_ga_breed via modules/genetic.py),
the evaluation is the fused CUDA path, and elitism, ranking and the curve statistics are one
more kernel: on a single GPU whole blocks of generations are enqueued on the device engine
(ggs_ga_run) and the host reads the state once per block.  Launched under torchrun the same
function shards the evaluation over the ranks (BASELINE config 4): every rank holds the whole
population and breeds the same next generation from the same counter-based random stream,
evaluates its contiguous slice, and one NCCL all-gather moves the fitness vector;
rank 0 alone writes frames and curves.
Two reference quirks are dropped because they cannot change results: elites are not
re-evaluated (the evaluation is deterministic, algorithm.py:134) -- their stored fitness is
reused -- and the offspring that elitism discards are not evaluated at all."""
import os
from typing import Tuple

import torch
import torch.nn.functional as F

try:
    from tqdm.auto import tqdm
except Exception:  # tqdm is optional
    def tqdm(it=None, **_):
        return it

from modules.fitness import fitness_many
from modules.genetic import breed_population
from modules.mask import compute_importance_mask
from modules.population import new_population
from ggs_b200.distributed import ShardedEvaluator, init_from_env, replicate
from ggs_b200.engine import MAX_POPULATION, GaEngine
from modules.utils import (_anneal_factor, build_mut_sigma, prewarm_renderer, save_curves_csv,
                           save_frame_png, save_loss_curve_png, scale_log_bounds)


def prepare_target(target_img_uint8: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """float32 [H,W,3] in [0,1], bilinearly resized to the work size (algorithm.py:33-39)."""
    t = target_img_uint8.to(torch.float32)
    if t.max() > 1.5:
        t = t / 255.0
    if t.shape[0] != H or t.shape[1] != W:
        t = F.interpolate(t.permute(2, 0, 1).unsqueeze(0), size=(H, W), mode='bilinear',
                          align_corners=False)[0].permute(1, 2, 0)
    return t.contiguous()


def _generations_on_engine(pop, target, imp_mask, H, W, k_sigma, boost_only, n_elite, generations,
                           run_seed, schedule, mut_sigma_max, mut_sigma_min, tour_k, cxpb, mutpb,
                           min_scale_splats, max_scale_splats, frame, progress, frame_every,
                           block=256, peers=None):
    """Generations are enqueued in blocks on the device engine (ggs_ga_run: breed, evaluate,
    elitism, ranking, curve point -- four launches per generation, no host sync); the host looks
    at the state once per block, and at every frame boundary.  With `peers` (one process per GPU)
    every rank runs this same loop: breeding and ranking are replicated, each rank evaluates its
    slice of the children and the raster kernel itself stores the fitness values into every
    rank's vector over NVLink -- still no host sync and no collective launch per generation."""
    P, N = int(pop.shape[0]), int(pop.shape[1])
    lo, hi = scale_log_bounds(H, W, min_scale_splats, max_scale_splats)
    eng = GaEngine(target, imp_mask, H, W, P, N, n_elite, generations, k_sigma=k_sigma,
                   boost_only=boost_only)
    try:
        if peers is not None:
            eng.set_peers(peers)
        eng.start(pop, run_seed)
        st = eng.state()
        frame(0, st["best_individual"])
        gen = 0
        try:
            while gen < generations:
                stop = min(generations, gen + block)
                if frame_every:
                    stop = min(stop, (gen // frame_every + 1) * frame_every)
                rows = [build_mut_sigma(g, generations, schedule, mut_sigma_max, mut_sigma_min)
                        for g in range(gen + 1, stop + 1)]
                eng.run(rows, tour_k, cxpb, mutpb, lo, hi)
                at_frame = bool(frame_every) and stop % frame_every == 0
                st = eng.state(curves_from=stop, want_best=at_frame)
                if at_frame:
                    frame(stop, st["best_individual"])
                progress(stop - gen, stop, st["best_fitness"], st["no_improve"])
                gen = stop
        except KeyboardInterrupt:
            print("\n[Interrupted] Returning current best individual...", flush=True)
        st = eng.state()
        c = st["curves"]
        curves = {"best": c[:, 0].tolist(), "mean": c[:, 1].tolist(), "median": c[:, 2].tolist()}
        return st["best_individual"], float(st["best_fitness"]), curves
    finally:
        eng.close()


def _generations_in_python(pop, evaluate, H, W, n_elite, generations, run_seed, schedule,
                           mut_sigma_max, mut_sigma_min, tour_k, cxpb, mutpb, min_scale_splats,
                           max_scale_splats, frame, progress, frame_every):
    """The same generations driven from Python: used when the evaluation is sharded over ranks
    (`evaluate` all-gathers the fitness vector) and for populations beyond the engine's limit.
    Same kernels, same counters: the result equals the engine's bit for bit."""
    pop_size, n_splats = int(pop.shape[0]), int(pop.shape[1])
    keep = pop_size - n_elite
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
    frame(0, best_ind)
    # Two generation buffers used alternately, each [n_elite + P, N, 9]: the breeding kernel writes
    # the children behind the elite rows, so the next population (elites first, then the first
    # `keep` children, algorithm.py:128-141) is a view of the buffer: no concatenation, no copy
    # of the children.
    room = [torch.empty((n_elite + pop_size, n_splats, 9), dtype=torch.float32, device=pop.device)
            for _ in range(2)]
    try:
        for gen in range(1, generations + 1):
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
            if gen % frame_every == 0:
                frame(gen, best_ind)
            progress(1, gen, best_fit, no_improve)
    except KeyboardInterrupt:
        print("\n[Interrupted] Returning current best individual...", flush=True)
    return best_ind.cpu(), best_fit, curves


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
    n_elite = max(1, elite_k)
    run_seed = int(replicate(torch.randint(0, 2**31 - 1, (1,)).to(device)).item())  # follows torch.manual_seed
    pad = len(str(generations))
    pbar = tqdm(total=generations, desc="GA generations", leave=True,
                **({"disable": True} if rank != 0 else {}))

    def frame(gen, individual):
        if save_video:
            save_frame_png(gen, individual, pad, prefix, video_dir, H, W, k_sigma, device, save_video)

    def progress(count, gen, best, stale):
        if hasattr(pbar, "update"):
            pbar.update(count)
            pbar.set_postfix(best_mse=f"{best:.6f}", stale=stale,
                             sigma_fac=f"{_anneal_factor(gen, generations, schedule):.3f}")

    use_engine = (pop_size <= MAX_POPULATION and os.environ.get("GGS_B200_GA_LOOP", "0") != "1")
    peers = None
    if use_engine and world > 1:
        # the engine shards its evaluation through peer-to-peer stores (ggs_b200.peers); if the
        # GPUs of this box cannot map each other's memory, fall back to the Python-driven loop
        # with one NCCL all-gather per generation
        try:
            from ggs_b200.peers import PeerGroup
            peers = PeerGroup.from_process_group(capacity=pop_size, device=target.device)
        except Exception as e:
            if rank == 0:
                print(f"[ggs_b200] peer-to-peer fitness exchange unavailable ({e}); using NCCL all-gather")
            use_engine = False
    try:
        if use_engine:
            best_ind, best_fit, curves = _generations_on_engine(
                pop, target, imp_mask, H, W, k_sigma, boost_only, n_elite, generations, run_seed,
                schedule, mut_sigma_max, mut_sigma_min, tour_k, cxpb, mutpb, min_scale_splats,
                max_scale_splats, frame, progress, max(1, frame_every) if save_video else 0,
                peers=peers)   # ranks may cut their blocks differently: epochs count generations
            if peers is not None:
                peers.check()
                import torch.distributed as dist
                dist.barrier()     # nobody unmaps a buffer another rank may still be writing to
                peers.close()
        else:
            best_ind, best_fit, curves = _generations_in_python(
                pop, evaluate, H, W, n_elite, generations, run_seed, schedule, mut_sigma_max,
                mut_sigma_min, tour_k, cxpb, mutpb, min_scale_splats, max_scale_splats, frame,
                progress, max(1, frame_every))
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
    return best_ind.cpu(), float(best_fit)
