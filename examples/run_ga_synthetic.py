#!/usr/bin/env python
"""Short GA / SA run on a synthetic target through the drop-in `modules` package.

    python examples/run_ga_synthetic.py [--generations 50] [--pop 64] [--splats 200] [--sa]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/run_ga_synthetic.py \
        --side 512 --splats 4000 --pop 8192 --generations 20        # BASELINE config 4, sharded

The reference's run_ggs.py needs imgs/reference.jpg, which it does not ship; this script builds
a target by rendering a hidden genome, so a perfect solution exists."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]

import torch  # noqa: E402

from modules import config as C  # noqa: E402
from modules.algorithm import genetic_approx  # noqa: E402
from modules.annealing import simulated_annealing  # noqa: E402
from modules.encode import genome_to_renderer_batched  # noqa: E402
from modules.population import new_population  # noqa: E402
from modules.render import render_splats_rgb_triton  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--generations", type=int, default=50)
    ap.add_argument("--pop", type=int, default=64)
    ap.add_argument("--splats", type=int, default=200)
    ap.add_argument("--side", type=int, default=128)
    ap.add_argument("--sa", action="store_true")
    a = ap.parse_args()
    torch.manual_seed(C.SEED)
    H = W = a.side
    hidden = new_population(1, a.splats, H, W, C.MIN_SCALE_SPLATS, C.MAX_SCALE_SPLATS, device="cuda")
    target = render_splats_rgb_triton(genome_to_renderer_batched(hidden), H, W, k_sigma=C.K_SIGMA,
                                      device="cuda")[0].cpu()
    common = dict(mut_sigma_max=C.MUT_SIGMA_MAX, mut_sigma_min=C.MUT_SIGMA_MIN,
                  min_scale_splats=C.MIN_SCALE_SPLATS, max_scale_splats=C.MAX_SCALE_SPLATS,
                  k_sigma=C.K_SIGMA, mask_strength=C.MASK_STRENGTH, boost_only=C.BOOST_ONLY)
    setup = 0.0
    if not a.sa:
        # Two 2-generation runs: the first loads the library / joins the process group /
        # allocates, the second measures the run's fixed cost (target, mask, initial population
        # and its evaluation) so the per-generation rate below is the marginal one.
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.time()
            genetic_approx(target, H, W, "cuda", pop_size=a.pop, n_splats=a.splats, generations=2,
                           tour_k=C.TOUR_K, elite_k=C.ELITE_K, cxpb=C.CXPB, mutpb=C.MUTPB,
                           schedule=C.SCHEDULE, **common)
            setup = time.time() - t0
    torch.cuda.synchronize()
    t0 = time.time()
    if a.sa:
        best, fit = simulated_annealing(target, H, W, "cuda", n_splats=a.splats, mutpb=C.MUTPB,
                                        sigma_schedule=C.SCHEDULE, iterations=a.generations,
                                        temp0=C.SA_T0, temp_schedule=C.SA_SCHEDULE,
                                        tries_per_iter=C.SA_TRIES_PER_ITER, **common)
    else:
        best, fit = genetic_approx(target, H, W, "cuda", pop_size=a.pop, n_splats=a.splats,
                                   generations=a.generations, tour_k=C.TOUR_K, elite_k=C.ELITE_K,
                                   cxpb=C.CXPB, mutpb=C.MUTPB, schedule=C.SCHEDULE, **common)
    dt = time.time() - t0
    if int(os.environ.get("RANK", "0")) == 0:
        per_gen = (dt - setup) / max(1, a.generations - 2) if not a.sa else dt / a.generations
        print(f"best fitness {fit:.6f} after {a.generations} generations in {dt:.2f} s; "
              f"{per_gen * 1e3:.2f} ms per generation ({1.0 / per_gen:.1f} per s) beyond the "
              f"{setup:.2f} s set-up, {int(os.environ.get('WORLD_SIZE', '1'))} process(es); "
              f"genome {tuple(best.shape)}")


if __name__ == "__main__":
    main()
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
