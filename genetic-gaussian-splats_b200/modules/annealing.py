"""Simulated annealing (reference: modules/annealing.py:29-190), same entry point.

Two schemes, both on the device engine (ggs_sa_run: no host in the loop; GGS_B200_SA_LOOP=1
drives the same chain from Python instead, same result bit for bit):

* batch_neighbors=False (DEFAULT) -- the reference's chain: `tries_per_iter` strictly sequential
  tries per iteration, each mutating the state the previous try left behind and evaluated on its
  own (B = 1, annealing.py:121-146).  Iterations are 1:1 with the reference's.
* batch_neighbors=True -- BASELINE config 2 "batched neighbour proposals": all tries of an
  iteration are proposed from the SAME state and scored by ONE evaluation, then the Metropolis
  test is applied to them in order.  About 4x the iterations per second, but NOT the reference's
  chain: a later try no longer starts from an earlier accepted try, so an iteration makes at most
  one effective move where the reference can make up to `tries_per_iter`; curves are not
  comparable iteration by iteration.

The scheme in use is printed once at the start of a run."""
from __future__ import annotations

import math
import os
import random
from typing import Tuple

import torch

try:
    from tqdm.auto import tqdm
except Exception:
    def tqdm(it, **_):
        return it

from modules.algorithm import prepare_target
from modules.fitness import fitness_many
from modules.mask import compute_importance_mask
from modules.population import new_individual
from ggs_b200 import breed
from ggs_b200.engine import MAX_TRIES, SaEngine
from modules.utils import (_anneal_factor, build_mut_sigma, prewarm_renderer, save_curves_csv,
                           save_frame_png, save_loss_curve_png, scale_log_bounds)

_ensure_hw = prepare_target  # the reference's private name (annealing.py:19-26)


def _temp_schedule(kind: str, T0: float, i: int, total: int) -> float:
    """Temperature at step i of `total` (annealing.py:29-44)."""
    n = max(1, total)
    p = i / n
    if kind == "linear":
        return max(1e-12, T0 * (1.0 - p))
    if kind == "cosine":
        return max(1e-12, T0 * 0.5 * (1.0 + math.cos(math.pi * p)))
    if kind == "log":
        return max(1e-12, T0 / (1.0 + math.log(1.0 + 9.0 * i)))
    if kind == "cauchy":
        return max(1e-12, T0 / (1.0 + i))
    return T0 * ((0.01 ** (1.0 / n)) ** i)   # "exp" and anything unknown


def _iterations_on_engine(curr, target, imp_mask, H, W, k_sigma, boost_only, iterations, run_seed,
                          temp0, temp_schedule, sigma_schedule, mut_sigma_max, mut_sigma_min, mutpb,
                          min_scale_splats, max_scale_splats, tries, batch_neighbors, frame,
                          frame_every, block=256):
    """Iterations enqueued in blocks on the device engine (ggs_sa_run: propose, evaluate,
    Metropolis -- four launches per try, or per iteration when the neighbours are batched; no
    host sync); the host looks at the state once per block, and at every frame boundary.  Same
    chain as _iterations_in_python with the same batch_neighbors, bit for bit."""
    lo, hi = scale_log_bounds(H, W, min_scale_splats, max_scale_splats)
    eng = SaEngine(target, imp_mask, H, W, int(curr.shape[0]), tries, iterations, k_sigma=k_sigma,
                   boost_only=boost_only, batch_neighbors=batch_neighbors)
    try:
        eng.start(curr, run_seed)
        frame(0, eng.state()["best_state"])
        pbar = tqdm(total=iterations, desc="SA iterations", leave=True)
        it = 0
        try:
            while it < iterations:
                stop = min(iterations, it + block)
                if frame_every:
                    stop = min(stop, (it // frame_every + 1) * frame_every)
                steps = range(it, stop)
                temps = [_temp_schedule(temp_schedule, temp0, i, iterations) for i in steps]
                eng.run([build_mut_sigma(i, iterations, sigma_schedule, mut_sigma_max, mut_sigma_min)
                         for i in steps], temps,
                        [[random.random() for _ in range(tries)] for _ in steps], mutpb, lo, hi)
                at_frame = bool(frame_every) and stop % frame_every == 0
                st = eng.state(curves_from=stop, want_states=at_frame)
                if at_frame:
                    frame(stop, st["best_state"])
                if hasattr(pbar, "update"):
                    pbar.update(stop - it)
                    pbar.set_postfix(best_mse=f"{st['best_energy']:.6f}",
                                     curr_mse=f"{st['current_energy']:.6f}", T=f"{temps[-1]:.4g}",
                                     sigma_fac=f"{_anneal_factor(stop - 1, iterations, sigma_schedule):.3f}")
                it = stop
        except KeyboardInterrupt:
            print("\n[Interrupted] Returning current best...", flush=True)
        finally:
            if hasattr(pbar, "close"):
                pbar.close()
        st = eng.state()
        c = st["curves"]
        return st["best_state"], float(st["best_energy"]), {"best": c[:, 0].tolist(),
                                                            "current": c[:, 1].tolist()}
    finally:
        eng.close()


def _iterations_in_python(curr, energy, propose, iterations, temp0, temp_schedule, sigma_schedule,
                          tries, batch_neighbors, frame, frame_every):
    """The loop of annealing.py:112-150 driven from Python: the reference's sequential tries
    (batch_neighbors=False), or the batched scheme one iteration at a time."""
    e_curr = float(energy(curr.unsqueeze(0))[0])
    best, best_fit = curr.clone(), e_curr
    curves = {"best": [best_fit], "current": [e_curr]}
    frame(0, best)
    pbar = tqdm(range(iterations), desc="SA iterations", leave=True)
    try:
        for it in pbar:
            T = _temp_schedule(temp_schedule, temp0, it, iterations)
            accepted_any = False
            uniforms = [random.random() for _ in range(tries)]  # one per try, used if uphill
            if batch_neighbors:
                neighbours = propose(curr, tries, it)
                energies = energy(neighbours)
            for k in range(tries):
                if batch_neighbors:
                    cand, e_new = neighbours[k], float(energies[k])
                else:
                    cand = propose(curr, 1, it)[0]
                    e_new = float(energy(cand.unsqueeze(0))[0])
                dE = e_new - e_curr
                if dE <= 0.0 or (T > 0.0 and uniforms[k] < math.exp(-dE / T)):
                    curr, e_curr, accepted_any = cand.clone(), e_new, True
                if e_curr + 1e-12 < best_fit:
                    best_fit, best = e_curr, curr.clone()
            curves["best"].append(best_fit)
            curves["current"].append(e_curr)
            if frame_every and (it + 1) % frame_every == 0:
                frame(it + 1, best)
            if hasattr(pbar, "set_postfix"):
                pbar.set_postfix(best_mse=f"{best_fit:.6f}", curr_mse=f"{e_curr:.6f}", T=f"{T:.4g}",
                                 accepted="Y" if accepted_any else "N",
                                 sigma_fac=f"{_anneal_factor(it, iterations, sigma_schedule):.3f}")
    except KeyboardInterrupt:
        print("\n[Interrupted] Returning current best...", flush=True)
    finally:
        if hasattr(pbar, "close"):
            pbar.close()
    return best, best_fit, curves


@torch.no_grad()
def simulated_annealing(
    target_img_uint8: torch.Tensor,
    H: int, W: int, device,
    n_splats: int,
    mutpb: float,
    mut_sigma_max: dict,
    mut_sigma_min: dict,
    sigma_schedule: str,
    min_scale_splats: float,
    max_scale_splats: float,
    k_sigma: float,
    mask_strength: float,
    boost_only: bool,
    iterations: int,
    temp0: float,
    temp_schedule: str,
    tries_per_iter: int = 1,
    save_video: bool = False,
    frame_every: int = 10_000,
    video_dir: str = "",
    prefix: str = "sa",
    loss_png_path: str = "",
    loss_csv_path: str = "",
    loss_log_y: bool = False,
    batch_neighbors: bool = False,
) -> Tuple[torch.Tensor, float]:
    """Minimises the (masked) MSE energy; returns (best axes-angle individual on CPU, best MSE).
    batch_neighbors=False is the reference's sequential chain; True evaluates the tries of an
    iteration in one launch (a different chain, see the module docstring)."""
    target = prepare_target(target_img_uint8, H, W).to(device)
    imp_mask = compute_importance_mask(target, H, W, edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3,
                                       gamma=0.7, floor=0.15, smooth=3, strength=mask_strength)
    prewarm_renderer(H, W, k_sigma, device)

    def energy(batch: torch.Tensor):
        return fitness_many(batch, target, H, W, k_sigma, device, tile=32, weight_mask=imp_mask,
                            boost_only=boost_only).cpu().tolist()

    run_seed = int(torch.randint(0, 2**31 - 1, (1,)).item())  # follows torch.manual_seed
    draws = [0]

    def propose(state: torch.Tensor, count: int, it: int) -> torch.Tensor:
        """`count` independently mutated copies of `state` (annealing.py:121-128, batched)."""
        nb = state.unsqueeze(0).repeat(count, 1, 1)
        assert nb.is_cuda, "simulated_annealing needs a CUDA device (no CPU path)"
        # one launch: the GA breeding kernel with identical parents and no crossover is
        # exactly `count` independent mutate_individual calls
        draws[0] += 1
        sigma = build_mut_sigma(it, iterations, sigma_schedule, mut_sigma_max, mut_sigma_min)
        lo, hi = scale_log_bounds(H, W, min_scale_splats, max_scale_splats)
        return breed(nb, torch.zeros(count, device=nb.device), sigma, tour_k=1, cxpb=0.0,
                     mutpb=mutpb, log_scale_lo=lo, log_scale_hi=hi, seed=run_seed,
                     generation=draws[0])

    curr = new_individual(n_splats, H, W, min_scale_splats, max_scale_splats, device=device)
    tries = max(1, tries_per_iter)
    pad = len(str(iterations))

    def frame(it: int, state: torch.Tensor) -> None:
        if save_video:
            save_frame_png(it, state.to(device), pad, prefix, video_dir, H, W, k_sigma, device,
                           save_video)

    print(f"[ggs_b200] simulated annealing, {tries} tries per iteration: " +
          ("batched neighbours -- all tries of an iteration start from the same state and are scored "
           "by one launch (NOT the reference's chain: at most one effective move per iteration)"
           if batch_neighbors and tries > 1 else
           "sequential tries, each from the state the previous one left (the reference's chain)"),
          flush=True)
    if curr.is_cuda and tries <= MAX_TRIES and os.environ.get("GGS_B200_SA_LOOP", "0") != "1":
        best, best_fit, curves = _iterations_on_engine(
            curr, target, imp_mask, H, W, k_sigma, boost_only, iterations, run_seed, temp0,
            temp_schedule, sigma_schedule, mut_sigma_max, mut_sigma_min, mutpb, min_scale_splats,
            max_scale_splats, tries, batch_neighbors, frame,
            max(1, frame_every) if save_video else 0)
    else:
        best, best_fit, curves = _iterations_in_python(
            curr, energy, propose, iterations, temp0, temp_schedule, sigma_schedule, tries,
            batch_neighbors, frame, max(1, frame_every) if save_video else 0)

    try:
        save_loss_curve_png(curves, loss_png_path, title=f"{prefix} energy (MSE)",
                            xlabel="Iteration", ylabel="MSE", log_y=loss_log_y, dpi=144)
        save_curves_csv(curves, loss_csv_path)
        if loss_png_path:
            print(f"Saved loss plot to {loss_png_path}")
        if loss_csv_path:
            print(f"Saved loss CSV to {loss_csv_path}")
    except Exception as e:
        print(f"[warn] Could not save SA curves: {e}")
    return best.cpu(), float(best_fit)
