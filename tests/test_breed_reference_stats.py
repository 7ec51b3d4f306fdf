"""The GA operators against statistics SAMPLED FROM THE REFERENCE's own genetic.py / population.py
(tests/golden/breed_reference_stats.npz, made by tests/golden/make_breed_golden.py): mutation
rates per gene group, noise standard deviations, the size-ordered swap (frequency, position,
direction), the "at least one gene per group" rule at mutpb = 0, the tournament winner
histogram, the row share of the uniform crossover, the initial-population quantiles.

CPU part: the torch restatement (oracle/torch_ref.py) and the host-side samplers
(modules/population.py, ggs_b200/synth.py).  GPU part: the one-launch breeding kernel
(ggs_ga_breed), whose counter-based streams differ from torch's, so only distributions can agree.
Tolerances are ~4 standard errors of the two sample sizes involved."""
import math
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = W = 128
LO, HI = math.log(3.0), math.log(12.8)


@pytest.fixture(scope="module")
def ref():
    with np.load(os.path.join(ROOT, "tests", "golden", "breed_reference_stats.npz")) as z:
        out = {k: z[k] for k in z.files}
    out["SIG"] = {str(k): float(v) for k, v in zip(out["sigma_names"], out["sigma_values"])}
    return out


COLS = {"x": 0, "y": 1, "alog": 2, "blog": 3, "theta": 4, "r": 5, "g": 6, "b": 7, "alpha": 8}
SIG_OF = {"x": "xy", "y": "xy", "alog": "alog", "blog": "blog", "theta": "theta", "r": "rgb",
          "g": "rgb", "b": "rgb", "alpha": "alpha"}


def mid_box_parent(N, seed=1):
    """The parent of make_breed_golden.py::parent(): clear of every clamp bound."""
    from ggs_b200 import synth
    g = synth.new_population_np(1, N, H, W, seed=seed)[0]
    g[:, 0:2] = 0.25 + 0.5 * g[:, 0:2]
    g[:, 2:4] = 0.5 * (LO + HI)
    g[:, 4] = 0.0
    g[:, 5:9] = 60.0 + 0.5 * g[:, 5:9]
    return torch.from_numpy(g)


def check_mutation_stats(ref, d):
    """d: [rows, 9] child - parent over rows outside the swap."""
    rows = d.shape[0]
    assert rows > 50_000
    for k, c in COLS.items():
        nz = d[:, c] != 0
        rate = float(nz.mean())
        se = math.sqrt(0.05 * 0.95 * (1.0 / rows + 1.0 / float(ref["mut_rows"])))
        assert abs(rate - float(ref[f"mut_rate_{k}"])) < 4.5 * se, (k, rate, float(ref[f"mut_rate_{k}"]))
        std = math.sqrt(float((d[nz, c].astype(np.float64) ** 2).mean())) / ref["SIG"][SIG_OF[k]]
        assert abs(std - float(ref[f"mut_std_over_sigma_{k}"])) < 0.045, (k, std)
    m = d[:, 5:8] != 0
    together = float((m.all(axis=1) | (~m).all(axis=1)).mean())
    assert together == float(ref["mut_rgb_all_or_none"]) == 1.0


def swap_summary(src, out):
    """src, out: [P, N, 9] numpy; noise off, so children are parents up to one row swap."""
    P, N = src.shape[:2]
    s0 = np.exp(src[..., 2] + src[..., 3])
    s1 = np.exp(out[..., 2] + out[..., 3])
    fired, i_sum, dist, fwd = 0, 0.0, 0.0, 0
    for p in range(P):
        moved = np.nonzero(np.abs(s0[p] - s1[p]) > 1e-6)[0]
        assert len(moved) in (0, 2)
        if len(moved):
            i, j = moved
            fired += 1
            i_sum += i / (N - 1)
            dist += (j - i) / N
            fwd += int(s1[p, i] > s1[p, j])
    return fired / P, i_sum / max(1, fired), dist / max(1, fired), fwd / max(1, fired)


def check_swap_stats(ref, freq, mean_i, mean_dist, fwd, samples):
    se = math.sqrt(0.9 * 0.1 * (1.0 / samples + 1.0 / float(ref["swap_reps"])))
    assert abs(freq - float(ref["swap_frequency"])) < 4.5 * se, freq
    assert abs(mean_i - float(ref["swap_mean_i_over_Nm1"])) < 0.025, mean_i
    assert abs(mean_dist - float(ref["swap_mean_distance_over_N"])) < 0.025, mean_dist
    assert fwd == float(ref["swap_bigger_first_fraction"]) == 1.0


def check_tournament_hist(ref, winners, k):
    want = ref[f"tour_hist_k{k}"]
    P = int(ref["tour_P"])
    got = np.bincount(winners, minlength=P) / len(winners)
    # both are samples of P(i) = ((P-i)^k - (P-i-1)^k) / P^k: compare with each other and with it
    i = np.arange(P)
    exact = ((P - i) ** k - (P - i - 1) ** k) / float(P ** k)
    assert 0.5 * np.abs(want - exact).sum() < 0.03          # the fixture itself
    assert 0.5 * np.abs(got - want).sum() < 0.045, 0.5 * np.abs(got - want).sum()
    assert abs((got * i).sum() - (want * i).sum()) < 0.6


# ------------------------------------------------------------------------------------- CPU

def test_torch_restatement_draws_from_the_reference_distributions(ref):
    from oracle import torch_ref as T
    N, reps = 400, 250
    src = mid_box_parent(N)
    gen = torch.Generator().manual_seed(5)
    pop = src.unsqueeze(0).repeat(reps, 1, 1)
    out = T.mutate_population(pop.clone(), 5, 10, "cosine", ref["SIG"], ref["SIG"], 0.05, H, W, 3.0,
                              0.1, generator=gen)
    d = (out - pop).numpy()
    stayed = (d != 0).sum(axis=-1) <= 5
    check_mutation_stats(ref, d[stayed])

    from ggs_b200 import synth
    zero = {k: 0.0 for k in ref["SIG"]}
    P, n = 6000, int(ref["swap_N"])
    rnd = torch.from_numpy(synth.new_population_np(P, n, H, W, seed=77))
    moved = T.mutate_population(rnd.clone(), 5, 10, "cosine", zero, zero, 0.0, H, W, 3.0, 0.1, generator=gen)
    check_swap_stats(ref, *swap_summary(rnd.numpy(), moved.numpy()), samples=P)

    fit = torch.arange(int(ref["tour_P"]), dtype=torch.float32)
    for k in (2, 3):
        check_tournament_hist(ref, T.tournament_indices(fit, 40000, k=k, generator=gen).numpy(), k)

    n = int(ref["cx_N"])
    parents = torch.stack([torch.zeros(n, 9), torch.ones(n, 9)] * 2000)
    kids = T.crossover_population(parents, cxpb=1.0, generator=gen)
    share = (kids[0::2, :, 0] == 0).float().mean(dim=1).numpy()
    assert abs(share.mean() - float(ref["cx_row_share_mean"])) < 0.006
    assert abs(share.std() - float(ref["cx_row_share_std"])) < 0.006
    assert bool((kids == kids[..., :1]).all()) and torch.equal(kids[0::2] + kids[1::2], torch.ones(2000, n, 9))


@pytest.mark.parametrize("h,w", [(128, 128), (200, 320)])
def test_initial_population_samplers_match_the_reference_quantiles(ref, h, w):
    from ggs_b200 import synth
    from modules.population import new_population
    q = ref["pop_quantile_levels"]
    torch.manual_seed(3)
    samplers = {"modules.population": new_population(40, 2000, h, w, 3.0, 0.1, device="cpu").numpy(),
                "synth": synth.new_population_np(40, 2000, h, w, seed=5)}
    tag = f"pop_{h}x{w}_"
    s_hi = 0.1 * max(h, w)
    for name, g in samplers.items():
        g = g.astype(np.float64)
        for key, vals, span in (("sigma_a", np.exp(g[..., 2]), s_hi - 3.0), ("sigma_b", np.exp(g[..., 3]), s_hi - 3.0),
                                ("xy", g[..., 0:2], 1.0), ("theta", g[..., 4], 2 * math.pi),
                                ("rgb", g[..., 5:8], 255.0), ("alpha", g[..., 8], 75.0)):
            err = np.abs(np.quantile(vals, q) - ref[tag + key]) / span
            assert err.max() < 0.012, (name, key, err)
        assert abs((g[..., 5:8] == 255.0).mean() - float(ref[tag + "rgb_at_255"])) < 0.002, name
        assert abs((g[..., 8] == 255.0).mean() - float(ref[tag + "alpha_at_255"])) < 0.004, name


# ------------------------------------------------------------------------------------- GPU

@pytest.fixture(scope="module")
def ggs():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ggs_b200
    ggs_b200.lib()
    return ggs_b200


def kernel_breed(ggs, pop, fit, sigma, **kw):
    args = dict(tour_k=2, cxpb=0.0, mutpb=0.05, log_scale_lo=LO, log_scale_hi=HI, seed=99, generation=3)
    args.update(kw)
    return ggs.breed(pop, fit, sigma, **args)


@pytest.mark.gpu
def test_kernel_mutation_rates_and_noise_match_the_reference(ggs, ref):
    N, reps = 400, 250
    src = mid_box_parent(N).cuda()
    pop = src.unsqueeze(0).repeat(reps, 1, 1).contiguous()      # identical parents: selection is moot
    out = kernel_breed(ggs, pop, torch.rand(reps, device="cuda"), ref["SIG"])
    d = (out - pop).cpu().numpy()
    stayed = (d != 0).sum(axis=-1) <= 5
    check_mutation_stats(ref, d[stayed])


@pytest.mark.gpu
def test_kernel_swap_matches_the_reference(ggs, ref):
    from ggs_b200 import synth
    zero = {k: 0.0 for k in ref["SIG"]}
    P, n = 2048, int(ref["swap_N"])
    g = synth.new_population_np(P, n, H, W, seed=77)
    g[..., 5] = 0.0
    g[..., 6] = (np.arange(P, dtype=np.float32) % 256)[:, None]            # parent tag, two bytes
    g[..., 7] = (np.arange(P, dtype=np.float32) // 256)[:, None]
    pop = torch.from_numpy(g).cuda()
    tot = np.zeros(4)
    gens = (3, 4, 5)
    for gen in gens:
        out = kernel_breed(ggs, pop, torch.rand(P, device="cuda"), zero, mutpb=0.0, generation=gen).cpu().numpy()
        parent = (out[:, 0, 6] + 256.0 * out[:, 0, 7]).round().astype(int)
        tot += np.asarray(swap_summary(g[parent], out))
    check_swap_stats(ref, *(tot / len(gens)), samples=P * len(gens))


@pytest.mark.gpu
def test_kernel_forced_genes_match_the_reference(ggs, ref):
    from ggs_b200 import synth
    n, P = int(ref["forced_N"]), 4000
    g = synth.new_population_np(1, n, H, W, seed=21)
    g[..., 0:2] = 0.25 + 0.5 * g[..., 0:2]
    g[..., 5:9] = 60.0 + 0.5 * g[..., 5:9]
    pop = torch.from_numpy(np.repeat(g, P, axis=0)).cuda()
    fit = torch.rand(P, device="cuda")
    zero = {k: 0.0 for k in ref["SIG"]}
    base = kernel_breed(ggs, pop, fit, zero, mutpb=0.0)
    noisy = kernel_breed(ggs, pop, fit, dict(ref["SIG"], alog=0.0, blog=0.0), mutpb=0.0)
    per_child = ((noisy - base).abs() > 1e-6).reshape(P, -1).sum(dim=1).cpu().numpy()
    hist = np.bincount(per_child, minlength=16)[:16] / P
    want = ref["forced_hist"]
    assert set(np.nonzero(hist)[0]) <= set(np.nonzero(want)[0]) == {3, 5}, hist
    assert np.abs(hist - want).max() < 4.5 * math.sqrt(0.25 * (1.0 / P + 1.0 / float(ref["forced_reps"])))
    scales = kernel_breed(ggs, pop, fit, dict(zero, alog=0.4, blog=0.4), mutpb=0.0)
    ia, ib = scales[..., 0].argsort(dim=1), base[..., 0].argsort(dim=1)
    sa = torch.gather(scales[..., 2:4], 1, ia.unsqueeze(-1).expand(-1, -1, 2))
    sb = torch.gather(base[..., 2:4], 1, ib.unsqueeze(-1).expand(-1, -1, 2))
    changed = ((sa - sb).abs() > 1e-6).reshape(P, -1).sum(dim=1).float().mean().item()
    # a forced scale gene can land on a clamp bound and stay put: never more than the reference's 1
    assert 0.9 < changed <= float(ref["forced_scale_genes_per_individual"]) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("k", [2, 3])
def test_kernel_tournament_histogram_matches_the_reference(ggs, ref, k):
    from ggs_b200 import synth
    P = int(ref["tour_P"])
    g = synth.new_population_np(P, 4, H, W, seed=2)
    g[..., 5] = np.arange(P, dtype=np.float32)[:, None]                      # parent tag
    pop = torch.from_numpy(g).cuda()
    fit = torch.arange(P, device="cuda", dtype=torch.float32)
    zero = {kk: 0.0 for kk in ref["SIG"]}
    winners = []
    for gen in range(1, 1 + 40000 // P):
        out = kernel_breed(ggs, pop, fit, zero, mutpb=0.0, tour_k=k, generation=gen)
        winners.append(out[:, 0, 5].round().long().cpu().numpy())
    check_tournament_hist(ref, np.concatenate(winners), k)


@pytest.mark.gpu
def test_kernel_crossover_row_share_matches_the_reference(ggs, ref):
    n, P = int(ref["cx_N"]), 4000
    pop = torch.zeros((P, n, 9), device="cuda")
    pop[:, :, 0:2] = 0.5
    pop[:, :, 2:4] = 0.5 * (LO + HI)
    pop[:, :, 5] = torch.arange(P, device="cuda", dtype=torch.float32)[:, None] % 251   # parent tag
    zero = {kk: 0.0 for kk in ref["SIG"]}
    out = kernel_breed(ggs, pop, torch.rand(P, device="cuda"), zero, mutpb=0.0, cxpb=1.0)
    # equal sizes: the swap never fires; rows of a pair come from its two parents, complementary
    a, b = out[0::2, :, 5], out[1::2, :, 5]
    first = a[:, :1]
    mixed = (a != b).all(dim=1) | (a == b).all(dim=1)
    assert bool(mixed.all())
    differs = (a != b).any(dim=1)                     # pairs whose two parents differ
    share = (a == first).float().mean(dim=1)[differs].cpu().numpy()
    # rows sharing the origin of row 0: 1 + Bin(n - 1, 1/2) of n, the reference's row share
    # Bin(n, 1/2) / n up to that one fixed row
    n_f = float(n)
    assert abs(share.mean() - (1.0 + 0.5 * (n_f - 1.0)) / n_f) < 0.006
    assert abs(share.std() - float(ref["cx_row_share_std"])) < 0.006
