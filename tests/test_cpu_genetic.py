"""CPU tests of the host-side callers of the hot path (modules/utils.py, population.py, ...) and of
the torch restatement of the GA operators (oracle/torch_ref.py) that the CUDA breeding kernel is
checked against."""
import math

import numpy as np
import torch

import modules.config as C
from modules import genetic as G
from oracle import torch_ref as T
from modules import population as P
from modules import resize as R
from modules import utils as U


def test_config_has_the_reference_constants():
    assert (C.N_SPLATS, C.POP_SIZE, C.K_SIGMA, C.DEFAULT_TILE_SIZE) == (512, 32, 3.0, 32)
    assert (C.MIN_SCALE_SPLATS, C.MAX_SCALE_SPLATS, C.ELITE_K, C.TOUR_K) == (3.0, 0.1, 8, 2)
    assert C.FRAME_EVERY == max(1, C.GENERATIONS // (C.FPS * C.VIDEO_LEN))
    assert set(C.MUT_SIGMA_MAX) == set(C.MUT_SIGMA_MIN) == {"xy", "alog", "blog", "theta", "rgb", "alpha"}


def test_new_population_ranges():
    torch.manual_seed(0)
    g = P.new_population(8, 300, 200, 120, 3.0, 0.1, device="cpu")
    assert g.shape == (8, 300, 9)
    assert g[..., :2].min() >= 0 and g[..., :2].max() <= 1
    s = g[..., 2:4].exp()
    assert s.min() >= 3.0 - 1e-4 and s.max() <= 20.0 + 1e-3
    assert g[..., 4].abs().max() <= math.pi
    assert g[..., 8].min() >= 180 and g[..., 5:9].max() <= 255
    assert len(P.population_to_list(g)) == 8 and P.new_individual(5, 8, 8, 3.0, 0.1, device="cpu").shape == (5, 9)


def test_schedules_and_clamp():
    assert U._anneal_factor(0, 100, "cosine") == 1.0 and abs(U._anneal_factor(100, 100, "cosine")) < 1e-12
    assert abs(U._anneal_factor(50, 100, "linear") - 0.5) < 1e-12
    assert abs(U._anneal_factor(100, 100, "exp") - 0.2) < 1e-9
    sig = U.build_mut_sigma(0, 10, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN)
    assert sig == C.MUT_SIGMA_MAX
    g = torch.tensor([[[-1.0, 2.0, -5.0, 9.0, 4.0, -3.0, 300.0, 128.0, 256.0]]])
    U.clamp_genome(g, 100, 50, 3.0, 0.1)
    assert g[0, 0, :2].tolist() == [0.0, 1.0]
    assert abs(g[0, 0, 2].item() - math.log(3.0)) < 1e-6 and abs(g[0, 0, 3].item() - math.log(10.0)) < 1e-6
    assert abs(g[0, 0, 4].item() - (4.0 - 2 * math.pi)) < 1e-6
    assert g[0, 0, 5:].tolist() == [0.0, 255.0, 128.0, 255.0]
    assert R.choose_work_size(1080, 1920, 512) == (288, 512) and R.choose_work_size(100, 50, 128) == (128, 64)
    ind = torch.zeros(3, 9)
    out = R.scale_genome_pixels_anisotropic(ind, 2.0, 4.0)
    assert torch.allclose(out[:, 2], torch.full((3,), math.log(4.0))) and ind.abs().sum() == 0


def test_mutation_keeps_genomes_legal_and_touches_every_group():
    torch.manual_seed(1)
    pop = P.new_population(16, 50, 64, 96, 3.0, 0.1, device="cpu")
    before = pop.clone()
    T.mutate_population(pop, 3, 100, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN, 0.0, 64, 96, 3.0, 0.1)
    lo, hi = U.scale_log_bounds(64, 96, 3.0, 0.1)
    assert pop[..., :2].min() >= 0 and pop[..., :2].max() <= 1
    assert pop[..., 2:4].min() >= lo - 1e-6 and pop[..., 2:4].max() <= hi + 1e-6
    assert pop[..., 5:9].min() >= 0 and pop[..., 5:9].max() <= 255
    # mutpb = 0: only the "at least one" genes and the swap change, yet every individual changes
    assert ((pop != before).reshape(16, -1).sum(1) > 0).all()


def test_swap_brings_a_bigger_splat_forward():
    torch.manual_seed(3)
    N = 30
    pop = P.new_population(64, N, 64, 64, 3.0, 0.1, device="cpu")
    size0 = (pop[..., 2] + pop[..., 3]).exp()
    T.mutate_population(pop, 100, 100, "cosine", {k: 0.0 for k in C.MUT_SIGMA_MAX},
                        {k: 0.0 for k in C.MUT_SIGMA_MIN}, 0.0, 64, 64, 3.0, 0.1)
    size1 = (pop[..., 2] + pop[..., 3]).exp()
    moved = (size0 - size1).abs() > 1e-6
    for p in range(64):
        idx = moved[p].nonzero().flatten().tolist()
        assert len(idx) in (0, 2)
        if idx:
            i, j = idx
            assert size1[p, i] > size1[p, j]          # the bigger one is now earlier
            assert torch.allclose(size0[p, i], size1[p, j]) and torch.allclose(size0[p, j], size1[p, i])


def test_tournament_and_crossover():
    torch.manual_seed(5)
    fit = torch.arange(100, dtype=torch.float32)
    idx = T.tournament_indices(fit, 4000, k=2)
    assert idx.min() >= 0 and idx.max() < 100
    assert idx.float().mean() < 40        # E[min of two uniform draws] ~ 33 < 49.5
    parents = torch.arange(6 * 4 * 9, dtype=torch.float32).reshape(6, 4, 9)
    same = T.crossover_population(parents, cxpb=0.0)
    assert torch.equal(same, parents)
    mixed = T.crossover_population(parents, cxpb=1.0)
    for k in range(3):       # rows are exchanged whole, pair-wise: the pair's multiset is kept
        pair_in = torch.cat([parents[2 * k], parents[2 * k + 1]]).sort(0).values
        pair_out = torch.cat([mixed[2 * k], mixed[2 * k + 1]]).sort(0).values
        assert torch.equal(pair_in, pair_out)
    a, b = parents[0], parents[1]
    c1, c2 = G.crossover_uniform(a, b)
    assert torch.equal(c1 + c2, a + b)
    pick = G.tournament_selection([a, b], [1.0, 0.0], k=8)
    assert pick.shape == a.shape


def test_temperature_schedule():
    from modules.annealing import _temp_schedule
    assert _temp_schedule("cosine", 1e-3, 0, 100) == 1e-3
    assert _temp_schedule("linear", 1.0, 100, 100) == 1e-12
    assert abs(_temp_schedule("exp", 1.0, 100, 100) - 0.01) < 1e-9
    assert _temp_schedule("cauchy", 1.0, 9, 100) == 0.1


def test_breed_population_torch_reference_keeps_shape_and_box():
    # the torch composition the CUDA breeding kernel is compared with: [P,N,9] legal genomes
    torch.manual_seed(2)
    pop = P.new_population(10, 30, 64, 64, 3.0, 0.1, device="cpu")
    fit = torch.rand(10)
    off = T.breed_population_torch(pop, fit, 3, 50, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN, 2,
                                   0.5, 0.05, 64, 64, 3.0, 0.1)
    lo, hi = U.scale_log_bounds(64, 64, 3.0, 0.1)
    assert off.shape == (10, 30, 9) and off.data_ptr() != pop.data_ptr()
    assert off[..., :2].min() >= 0 and off[..., :2].max() <= 1
    assert off[..., 2:4].min() >= lo - 1e-6 and off[..., 2:4].max() <= hi + 1e-6
    assert off[..., 5:9].min() >= 0 and off[..., 5:9].max() <= 255


def test_product_operators_have_no_cpu_path():
    # modules/genetic.py and modules/mask.py are device-only: a CPU box fails loudly
    import pytest
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    pop = P.new_population(4, 6, 32, 32, 3.0, 0.1, device="cpu")
    with pytest.raises(AssertionError):
        G.breed_population(pop, torch.rand(4), 1, 10, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN, 2,
                           0.5, 0.05, 32, 32, 3.0, 0.1)
    with pytest.raises(AssertionError):
        G.mutate_individual(pop[0], False, 1, 10, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN, 0.05,
                            32, 32, 3.0, 0.1)
    from modules.mask import compute_importance_mask
    with pytest.raises(AssertionError):
        compute_importance_mask(torch.rand(16, 16, 3), 16, 16)


def test_schedules_agree_with_the_torch_reference():
    for kind in ("cosine", "linear", "exp", "other"):
        for g in (0, 1, 37, 100, 150):
            assert abs(U._anneal_factor(g, 100, kind) - T.anneal_factor(g, 100, kind)) < 1e-15
    assert U.build_mut_sigma(30, 100, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN) == \
        T.build_mut_sigma(30, 100, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN)
