#!/usr/bin/env python
"""Distribution fixtures for the GA operators, SAMPLED FROM THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference).  The reference's operators are
per-individual torch code (CPU tensors work: every op uses `ind.device`):

  modules.genetic.mutate_individual      (genetic.py:32-92)
  modules.genetic.tournament_selection   (genetic.py:8-14)
  modules.genetic.crossover_uniform      (genetic.py:17-21)
  modules.population.new_population      (population.py:20-46)

They are called a few 10^4 times on seeded inputs and only STATISTICS of their outputs are
saved (tests/golden/breed_reference_stats.npz): mutation rates per gene group, noise standard
deviations, how often and where the size-ordered swap fires, the number of genes that change
when mutpb = 0 (the "at least one" rule), the tournament winner histogram, the row share of the
uniform crossover, and quantiles of the initial-population distribution.  The CUDA breeding
kernel (ggs_ga_breed) draws from its own counter-based streams, so it is compared with these
statistics, not with samples (tests/test_gpu_breed.py).

Nothing of the reference is copied into the repo: only numbers.

Usage:  python tests/golden/make_breed_golden.py
"""
import math
import os
import random
import sys

REF = os.environ.get("GGS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import modules.genetic as rgen  # noqa: E402
import modules.population as rpop  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# the test population of tests/test_gpu_breed.py::population(): clear of every clamp bound
H = W = 128
SIG = {"xy": 0.05, "alog": 0.3, "blog": 0.2, "theta": 0.1, "rgb": 10.0, "alpha": 5.0}
ZERO = {k: 0.0 for k in SIG}
LO, HI = math.log(3.0), math.log(0.1 * 128)


def parent(N, seed):
    torch.manual_seed(seed)
    g = rpop.new_population(1, N, H, W, 3.0, 0.1, device="cpu")[0]
    g[:, 0:2] = 0.25 + 0.5 * g[:, 0:2]
    g[:, 2:4] = 0.5 * (LO + HI)
    g[:, 4] = 0.0
    g[:, 5:9] = 60.0 + 0.5 * g[:, 5:9]
    return g


def mutate(ind, sig, mutpb):
    # mut_sigma_max == mut_sigma_min: the schedule cannot matter
    return rgen.mutate_individual(ind.clone(), False, 5, 10, "cosine", sig, sig, mutpb, H, W, 3.0, 0.1)


def mutation_stats(N=400, reps=250, mutpb=0.05):
    """Rates and noise std per gene group over rows that did not take part in the swap."""
    src = parent(N, 1)
    cols = {"x": 0, "y": 1, "alog": 2, "blog": 3, "theta": 4, "r": 5, "g": 6, "b": 7, "alpha": 8}
    sig_of = {"x": "xy", "y": "xy", "alog": "alog", "blog": "blog", "theta": "theta", "r": "rgb",
              "g": "rgb", "b": "rgb", "alpha": "alpha"}
    changed = {k: 0 for k in cols}
    sq = {k: 0.0 for k in cols}
    rows = 0
    rgb_together = rgb_rows = 0
    for r in range(reps):
        torch.manual_seed(1000 + r)
        d = (mutate(src, SIG, mutpb) - src).numpy()
        stayed = (d != 0).sum(axis=1) <= 5          # a swapped row differs in x, y, r, g, b, alpha
        d = d[stayed]
        rows += len(d)
        for k, c in cols.items():
            nz = d[:, c] != 0
            changed[k] += int(nz.sum())
            sq[k] += float((d[nz, c].astype(np.float64) ** 2).sum())
        m = d[:, 5:8] != 0
        rgb_rows += len(d)
        rgb_together += int((m.all(axis=1) | (~m).all(axis=1)).sum())
    out = {"mut_rows": rows, "mut_mutpb": mutpb, "mut_rgb_all_or_none": rgb_together / rgb_rows}
    for k in cols:
        out[f"mut_rate_{k}"] = changed[k] / rows
        out[f"mut_std_over_sigma_{k}"] = math.sqrt(sq[k] / max(1, changed[k])) / SIG[sig_of[k]]
    return out


def swap_stats(N=30, reps=6000):
    """The size-ordered swap with noise switched off: frequency, position of i, distance j - i."""
    fired, i_sum, dist_sum, forward = 0, 0.0, 0.0, 0
    for r in range(reps):
        torch.manual_seed(20000 + r)
        src = rpop.new_population(1, N, H, W, 3.0, 0.1, device="cpu")[0]
        out = mutate(src, ZERO, 0.0)
        size0 = (src[:, 2] + src[:, 3]).exp()
        size1 = (out[:, 2] + out[:, 3]).exp()
        moved = ((size0 - size1).abs() > 1e-6).nonzero().flatten().tolist()
        if moved:
            i, j = moved
            fired += 1
            i_sum += i / (N - 1)
            dist_sum += (j - i) / N
            forward += int(size1[i] > size1[j])       # the bigger splat now sits earlier
    return {"swap_N": N, "swap_reps": reps, "swap_frequency": fired / reps,
            "swap_mean_i_over_Nm1": i_sum / max(1, fired), "swap_mean_distance_over_N": dist_sum / max(1, fired),
            "swap_bigger_first_fraction": forward / max(1, fired)}


def forced_gene_stats(N=20, reps=4000):
    """mutpb = 0: only the forced genes change.  Histogram of changed genes per individual, with
    the scale noise off (rows line up: the swap does not depend on noise then)."""
    sig = dict(SIG, alog=0.0, blog=0.0)
    hist = np.zeros(16, dtype=np.int64)
    scale_changed = 0
    for r in range(reps):
        torch.manual_seed(40000 + r)
        src = rpop.new_population(1, N, H, W, 3.0, 0.1, device="cpu")[0]
        src[:, 0:2] = 0.25 + 0.5 * src[:, 0:2]
        src[:, 5:9] = 60.0 + 0.5 * src[:, 5:9]
        torch.manual_seed(50000 + r)
        a = mutate(src, sig, 0.0)
        torch.manual_seed(50000 + r)
        b = mutate(src, ZERO, 0.0)                    # same draws, no noise: the swap alone
        hist[int(((a - b).abs() > 1e-6).sum())] += 1
        torch.manual_seed(50000 + r)
        c = mutate(src, dict(ZERO, alog=0.4, blog=0.4), 0.0)
        ia, ib = c[:, 0].argsort(), b[:, 0].argsort()
        scale_changed += int(((c[ia, 2:4] - b[ib, 2:4]).abs() > 1e-6).sum())
    return {"forced_N": N, "forced_reps": reps, "forced_hist": hist / reps,
            "forced_scale_genes_per_individual": scale_changed / reps}


def tournament_stats(P=64, draws=40000):
    pop = [torch.full((1, 9), float(i)) for i in range(P)]
    fits = [float(i) for i in range(P)]
    out = {}
    for k in (2, 3):
        random.seed(7 + k)
        h = np.zeros(P, dtype=np.int64)
        for _ in range(draws):
            h[int(rgen.tournament_selection(pop, fits, k=k)[0, 0])] += 1
        out[f"tour_hist_k{k}"] = h / draws
    out["tour_P"] = P
    out["tour_draws"] = draws
    return out


def crossover_stats(N=64, reps=4000):
    torch.manual_seed(9)
    a, b = torch.zeros(N, 9), torch.ones(N, 9)
    share, whole, complementary = [], 0, 0
    for _ in range(reps):
        c1, c2 = rgen.crossover_uniform(a, b)
        share.append(float((c1[:, 0] == 0).float().mean()))
        whole += int(((c1 == c1[:, :1]).all()) and ((c2 == c2[:, :1]).all()))
        complementary += int(torch.equal(c1 + c2, a + b))
    share = np.asarray(share)
    return {"cx_N": N, "cx_reps": reps, "cx_row_share_mean": float(share.mean()),
            "cx_row_share_std": float(share.std()), "cx_rows_whole": whole / reps,
            "cx_complementary": complementary / reps}


def population_stats(B=40, N=2000):
    torch.manual_seed(11)
    q = np.array([0.01, 0.05, 0.25, 0.5, 0.75, 0.95, 0.99])
    out = {"pop_quantile_levels": q}
    for (h, w) in ((128, 128), (200, 320)):
        g = rpop.new_population(B, N, h, w, 3.0, 0.1, device="cpu").numpy().astype(np.float64)
        tag = f"pop_{h}x{w}_"
        out[tag + "sigma_a"] = np.quantile(np.exp(g[..., 2]), q)
        out[tag + "sigma_b"] = np.quantile(np.exp(g[..., 3]), q)
        out[tag + "xy"] = np.quantile(g[..., 0:2], q)
        out[tag + "theta"] = np.quantile(g[..., 4], q)
        out[tag + "rgb"] = np.quantile(g[..., 5:8], q)
        out[tag + "alpha"] = np.quantile(g[..., 8], q)
        out[tag + "rgb_at_255"] = float((g[..., 5:8] == 255.0).mean())
        out[tag + "alpha_at_255"] = float((g[..., 8] == 255.0).mean())
    return out


def main():
    stats = {}
    for part in (mutation_stats, swap_stats, forced_gene_stats, tournament_stats, crossover_stats,
                 population_stats):
        got = part()
        stats.update(got)
        print(part.__name__, {k: (np.round(v, 4).tolist() if isinstance(v, np.ndarray) else round(v, 5)
                                  if isinstance(v, float) else v) for k, v in got.items()}, flush=True)
    stats["sigma_names"] = np.array(sorted(SIG))
    stats["sigma_values"] = np.array([SIG[k] for k in sorted(SIG)])
    np.savez_compressed(os.path.join(HERE, "breed_reference_stats.npz"), **stats)
    print("wrote breed_reference_stats.npz with", len(stats), "entries")


if __name__ == "__main__":
    main()
