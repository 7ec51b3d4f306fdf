"""GPU tests of the on-device GA engine (ggs_ga_*): elitism, ranking, curve statistics and
best-so-far bookkeeping (algorithm.py:128-160) without the host in the loop."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ggs():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ggs_b200
    ggs_b200.lib()
    return ggs_b200


def setup(P, N, H, W, seed=0):
    from ggs_b200 import synth
    pop = torch.from_numpy(synth.new_population_np(P, N, H, W, seed=seed)).cuda()
    t_np = synth.synthetic_target_np(H, W, seed)
    return pop, torch.from_numpy(t_np).cuda(), torch.from_numpy(synth.importance_mask_np(t_np)).cuda()


SIG = {"xy": 0.05, "alog": 0.3, "blog": 0.2, "theta": 0.1, "rgb": 10.0, "alpha": 5.0}
LO, HI = math.log(3.0), math.log(9.6)


def host_generation(ggs, pop, fit, order, t, m, H, W, n_elite, gen, seed):
    """One generation with the public device ops and torch for the bookkeeping."""
    P = pop.shape[0]
    keep = P - n_elite
    children = ggs.breed(pop, fit, SIG, tour_k=2, cxpb=0.3, mutpb=0.1, log_scale_lo=LO,
                         log_scale_hi=HI, seed=seed, generation=gen)
    child_fit = ggs.fitness(children[:keep].contiguous(), t, H, W, 3.0, weight_mask=m)
    new_pop = torch.cat([pop[order[:n_elite]], children[:keep]])
    new_fit = torch.cat([fit[order[:n_elite]], child_fit])
    return new_pop, new_fit, torch.argsort(new_fit, stable=True)


@pytest.mark.parametrize("P,N,n_elite", [(32, 60, 8), (7, 33, 1), (50, 20, 0), (6, 10, 6), (300, 12, 5)])
def test_engine_generations_match_host_bookkeeping(ggs, P, N, n_elite):
    from ggs_b200.engine import GaEngine
    H, W, G, seed = 48, 64, 9, 77
    pop, t, m = setup(P, N, H, W, seed=P)
    eng = GaEngine(t, m, H, W, P, N, n_elite, G)
    eng.start(pop, seed)
    fit = ggs.fitness(pop, t, H, W, 3.0, weight_mask=m)
    order = torch.argsort(fit, stable=True)
    best_fit, best_ind, stale = float(fit[order[0]]), pop[order[0]].clone(), 0
    want = [(best_fit, float(fit.double().mean()), float(fit.double().median()) if P % 2 else None)]
    st = eng.state()
    assert st["generation"] == 0 and st["best_fitness"] == best_fit and st["no_improve"] == 0
    assert torch.equal(st["best_individual"], best_ind.cpu())
    for gen in range(1, G + 1):
        eng.run([SIG], 2, 0.3, 0.1, LO, HI)
        pop, fit, order = host_generation(ggs, pop, fit, order, t, m, H, W, n_elite, gen, seed)
        gbest = float(fit[order[0]])
        if gbest + 1e-10 < best_fit:
            best_fit, best_ind, stale = gbest, pop[order[0]].clone(), 0
        else:
            stale += 1
        e_pop, e_fit = eng.population()
        assert torch.equal(e_pop, pop), gen
        assert torch.equal(e_fit, fit), gen
        st = eng.state(curves_from=gen)
        assert st["generation"] == gen and st["no_improve"] == stale
        assert st["best_fitness"] == best_fit
        assert torch.equal(st["best_individual"], best_ind.cpu())
        ranked = fit[order].double()
        c = st["curves"][0]
        assert c[0] == best_fit
        assert abs(c[1] - float(ranked.mean())) <= 1e-12 * max(1.0, abs(c[1]))
        assert c[2] == float(0.5 * (ranked[(P - 1) // 2] + ranked[P // 2]))
    assert eng.state()["curves"].shape == (G + 1, 3)
    eng.close()


def test_engine_blocks_equal_single_steps(ggs):
    """Enqueuing 12 generations in one call, or in 12 calls, is the same run."""
    from ggs_b200.engine import GaEngine
    P, N, H, W, G = 24, 40, 64, 64, 12
    pop, t, m = setup(P, N, H, W, seed=5)
    rows = [{k: v * (1.0 - 0.05 * g) for k, v in SIG.items()} for g in range(G)]
    out = []
    for block in (G, 1, 5):
        eng = GaEngine(t, m, H, W, P, N, 4, G, boost_only=(block == 5) and False)
        eng.start(pop, 9)
        for g0 in range(0, G, block):
            eng.run(rows[g0:g0 + block], 2, 0.2, 0.1, LO, HI)
        st = eng.state()
        out.append((st["best_fitness"], st["best_individual"], st["curves"], eng.population()[0]))
        eng.close()
    for other in out[1:]:
        assert other[0] == out[0][0] and torch.equal(other[1], out[0][1])
        assert np.array_equal(other[2], out[0][2]) and torch.equal(other[3], out[0][3])


def test_engine_modes_and_plain_fitness(ggs):
    from ggs_b200.engine import GaEngine
    P, N, H, W = 10, 30, 40, 56
    pop, t, m = setup(P, N, H, W, seed=2)
    for mask, boost in ((None, False), (m, False), (m, True)):
        eng = GaEngine(t, mask, H, W, P, N, 2, 3, boost_only=boost)
        eng.start(pop, 1)
        _, fit = eng.population()
        want = ggs.fitness(pop, t, H, W, 3.0, weight_mask=mask, boost_only=boost)
        assert torch.equal(fit.sort().values, want.sort().values)
        eng.close()


def test_engine_errors(ggs):
    from ggs_b200.engine import GaEngine
    pop, t, m = setup(8, 10, 32, 32)
    with pytest.raises(ggs.GgsError):
        GaEngine(t, m, 32, 32, 20000, 10, 1, 5)           # population beyond the ranking limit
    with pytest.raises(ggs.GgsError):
        GaEngine(t, m, 32, 32, 8, 10, 9, 5)               # more elites than individuals
    eng = GaEngine(t, m, 32, 32, 8, 10, 2, 3)
    with pytest.raises(ggs.GgsError):
        eng.run([SIG], 2, 0.1, 0.1, LO, HI)               # not started
    eng.start(pop, 3)
    eng.run([SIG] * 3, 2, 0.1, 0.1, LO, HI)
    with pytest.raises(ggs.GgsError):
        eng.run([SIG], 2, 0.1, 0.1, LO, HI)               # beyond max_generations
    assert eng.state()["generation"] == 3
    eng.close()


def test_genetic_approx_engine_equals_python_loop(ggs, tmp_path):
    """The GA entry point gives the same run on the engine and on the Python-driven loop."""
    import modules.config as C
    from ggs_b200 import synth
    from modules.algorithm import genetic_approx
    H, W = 48, 72
    target = torch.from_numpy(synth.synthetic_target_np(H, W, 11))
    kw = dict(H=H, W=W, device="cuda", pop_size=18, n_splats=35, generations=40, tour_k=C.TOUR_K,
              elite_k=3, cxpb=0.3, mutpb=C.MUTPB, mut_sigma_max=C.MUT_SIGMA_MAX,
              mut_sigma_min=C.MUT_SIGMA_MIN, schedule=C.SCHEDULE, min_scale_splats=C.MIN_SCALE_SPLATS,
              max_scale_splats=C.MAX_SCALE_SPLATS, k_sigma=C.K_SIGMA, mask_strength=C.MASK_STRENGTH,
              boost_only=C.BOOST_ONLY)
    runs = []
    for loop, frames in (("0", False), ("1", False), ("0", True)):
        os.environ["GGS_B200_GA_LOOP"] = loop
        os.environ["TQDM_DISABLE"] = "1"
        try:
            torch.manual_seed(3)
            extra = {}
            if frames:
                d = tmp_path / "frames"
                d.mkdir()
                extra = dict(save_video=True, frame_every=16, video_dir=str(d), prefix="t",
                             loss_csv_path=str(tmp_path / "loss.csv"))
            runs.append(genetic_approx(target, **kw, **extra))
        finally:
            os.environ.pop("GGS_B200_GA_LOOP", None)
    for best, fit in runs[1:]:
        assert fit == runs[0][1] and torch.equal(best, runs[0][0])
    # frames at generation 0, 16, 32 and the curve file with one line per generation
    assert len(list((tmp_path / "frames").glob("*.png"))) == 3
    assert len((tmp_path / "loss.csv").read_text().strip().splitlines()) == 40 + 2


# ---- simulated annealing engine (ggs_sa_*) -----------------------------------------------------

def host_sa_iteration(ggs, cur, e_cur, best, e_best, t, m, H, W, tries, it, seed, T, uniforms,
                      batched=True):
    """One iteration with the public device ops; Metropolis on the host (annealing.py:121-146):
    batched (all neighbours from the same state, one evaluation) or the reference's sequential
    tries (each from the state the previous one left, proposal number keys the random stream)."""
    if batched:
        nb = cur.unsqueeze(0).repeat(tries, 1, 1).contiguous()
        cand = ggs.breed(nb, torch.zeros(tries, device="cuda"), SIG, tour_k=1, cxpb=0.0, mutpb=0.1,
                         log_scale_lo=LO, log_scale_hi=HI, seed=seed, generation=it)
        en = ggs.fitness(cand, t, H, W, 3.0, weight_mask=m).cpu().tolist()
    for k in range(tries):
        if not batched:
            one = ggs.breed(cur.unsqueeze(0).contiguous(), torch.zeros(1, device="cuda"), SIG, tour_k=1,
                            cxpb=0.0, mutpb=0.1, log_scale_lo=LO, log_scale_hi=HI, seed=seed,
                            generation=(it - 1) * tries + k + 1)
            cand = {k: one[0]}
            en = {k: float(ggs.fitness(one, t, H, W, 3.0, weight_mask=m)[0])}
        dE = en[k] - e_cur
        if dE <= 0.0 or (T > 0.0 and uniforms[k] < math.exp(-dE / T)):
            cur, e_cur = cand[k].clone(), en[k]
        if e_cur + 1e-12 < e_best:
            e_best, best = e_cur, cur.clone()
    return cur, e_cur, best, e_best


@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("N,tries,T0", [(40, 8, 2e-3), (33, 1, 1e-3), (12, 64, 5e-3), (25, 5, 0.0),
                                        (1, 3, 1e-3),       # one splat: no swap, every group forced
                                        (4200, 2, 1e-3)])   # beyond the proposal kernel: breed + decode
def test_sa_engine_iterations_match_host_metropolis(ggs, N, tries, T0, batched):
    from ggs_b200.engine import SaEngine
    H, W, I, seed = 48, 64, (14 if batched or tries < 64 else 4) if N < 1000 else 3, 31
    pop, t, m = setup(1, N, H, W, seed=N)
    rng = np.random.default_rng(tries)
    eng = SaEngine(t, m, H, W, N, tries, I, batch_neighbors=batched)
    eng.start(pop[0], seed)
    cur = pop[0].clone()
    e_cur = float(ggs.fitness(pop, t, H, W, 3.0, weight_mask=m)[0])
    best, e_best = cur.clone(), e_cur
    st = eng.state()
    assert st["iteration"] == 0 and st["best_energy"] == e_cur and st["current_energy"] == e_cur
    assert torch.equal(st["best_state"], cur.cpu()) and torch.equal(st["current_state"], cur.cpu())
    uphill_accepts = 0
    for it in range(1, I + 1):
        T = T0 * (1.0 - it / (I + 1))
        u = rng.random(tries).tolist()
        eng.run([SIG], [T], [u], 0.1, LO, HI)
        before = e_cur
        cur, e_cur, best, e_best = host_sa_iteration(ggs, cur, e_cur, best, e_best, t, m, H, W, tries,
                                                     it, seed, T, u, batched=batched)
        uphill_accepts += e_cur > before
        st = eng.state(curves_from=it)
        assert st["iteration"] == it
        assert st["current_energy"] == e_cur and st["best_energy"] == e_best, it
        assert torch.equal(st["current_state"], cur.cpu()), it
        assert torch.equal(st["best_state"], best.cpu()), it
        assert st["curves"].tolist() == [[e_best, e_cur]]
    if T0 == 0.0:
        assert uphill_accepts == 0
    assert eng.state()["curves"].shape == (I + 1, 2)
    eng.close()


def test_sa_engine_blocks_equal_single_steps(ggs):
    from ggs_b200.engine import SaEngine
    N, H, W, I, tries = 50, 64, 64, 12, 6
    pop, t, m = setup(1, N, H, W, seed=4)
    rows = [{k: v * (1.0 - 0.05 * g) for k, v in SIG.items()} for g in range(I)]
    temps = [3e-3 * (1.0 - g / I) for g in range(I)]
    uni = np.random.default_rng(0).random((I, tries)).tolist()
    out = []
    for block, batched in ((I, True), (1, True), (5, True), (I, False), (4, False)):
        eng = SaEngine(t, None if block == 5 else m, H, W, N, tries, I, batch_neighbors=batched)
        eng.start(pop[0], 9)
        for g0 in range(0, I, block):
            eng.run(rows[g0:g0 + block], temps[g0:g0 + block], uni[g0:g0 + block], 0.1, LO, HI)
        st = eng.state()
        out.append((st["best_energy"], st["best_state"], st["curves"], st["current_state"]))
        eng.close()
    assert out[1][0] == out[0][0] and torch.equal(out[1][1], out[0][1])
    assert np.array_equal(out[1][2], out[0][2]) and torch.equal(out[1][3], out[0][3])
    assert out[2][0] != out[0][0]          # the plain (unmasked) energy is a different run
    # sequential tries: blocks equal blocks, and the chain differs from the batched one
    assert out[4][0] == out[3][0] and torch.equal(out[4][1], out[3][1])
    assert np.array_equal(out[4][2], out[3][2]) and torch.equal(out[4][3], out[3][3])
    assert not np.array_equal(out[3][2], out[0][2])


def test_sa_engine_errors(ggs):
    from ggs_b200.engine import SaEngine
    pop, t, m = setup(1, 10, 32, 32)
    with pytest.raises(ggs.GgsError):
        SaEngine(t, m, 32, 32, 10, 65, 5)                 # more tries than the kernel takes
    eng = SaEngine(t, m, 32, 32, 10, 2, 3)
    with pytest.raises(ggs.GgsError):
        eng.run([SIG], [1e-3], [[0.5, 0.5]], 0.1, LO, HI)  # not started
    eng.start(pop[0], 3)
    eng.run([SIG] * 3, [1e-3] * 3, [[0.5, 0.5]] * 3, 0.1, LO, HI)
    with pytest.raises(ggs.GgsError):
        eng.run([SIG], [1e-3], [[0.5, 0.5]], 0.1, LO, HI)  # beyond max_iterations
    assert eng.state()["iteration"] == 3
    eng.close()


def test_simulated_annealing_engine_equals_python_loop(ggs, tmp_path):
    """The SA entry point gives the same chain on the engine and on the Python-driven loop."""
    import modules.config as C
    from ggs_b200 import synth
    from modules.annealing import simulated_annealing
    import random
    H, W = 48, 72
    target = torch.from_numpy(synth.synthetic_target_np(H, W, 11))
    kw = dict(H=H, W=W, device="cuda", n_splats=35, mutpb=0.05, mut_sigma_max=C.MUT_SIGMA_MAX,
              mut_sigma_min=C.MUT_SIGMA_MIN, sigma_schedule="cosine", min_scale_splats=3.0,
              max_scale_splats=0.1, k_sigma=3.0, mask_strength=0.7, boost_only=False, iterations=45,
              temp0=2e-3, temp_schedule="cosine", tries_per_iter=8)
    for batched in (False, True):
        _sa_engine_vs_loop(target, dict(kw, batch_neighbors=batched), tmp_path / f"b{int(batched)}")


def _sa_engine_vs_loop(target, kw, tmp_path):
    from modules.annealing import simulated_annealing
    import random
    tmp_path.mkdir()
    runs = []
    for loop, frames in (("0", False), ("1", False), ("0", True)):
        os.environ["GGS_B200_SA_LOOP"] = loop
        os.environ["TQDM_DISABLE"] = "1"
        try:
            torch.manual_seed(3)
            random.seed(5)
            extra = {}
            if frames:
                d = tmp_path / "frames"
                d.mkdir()
                extra = dict(save_video=True, frame_every=16, video_dir=str(d), prefix="t",
                             loss_csv_path=str(tmp_path / "loss.csv"))
            runs.append(simulated_annealing(target, **kw, **extra))
        finally:
            os.environ.pop("GGS_B200_SA_LOOP", None)
    for best, e in runs[1:]:
        assert e == runs[0][1] and torch.equal(best, runs[0][0])
    assert len(list((tmp_path / "frames").glob("*.png"))) == 3   # iteration 0, 16, 32
    assert len((tmp_path / "loss.csv").read_text().strip().splitlines()) == 45 + 2
