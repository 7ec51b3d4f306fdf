"""The peer-to-peer fitness exchange (ggs_peers_*, ggs_fitness_allgather, ggs_ga_set_peers) on ONE
GPU: several ranks live in this process (ggs_peers_connect_local), each with its own stream, so
the whole protocol -- stores into every rank's vector from inside the raster kernel, the
system-scope flags, the waits, the two-epoch buffering, the sharded GA engine -- runs wherever
the driver runs the GPU suite.  The one-process-per-GPU form over CUDA IPC is covered by the
2-GPU tests in test_gpu_search.py and by bench.py --gpus N (`gather_bit_identical`)."""
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


def inputs(P, N, H, W, seed):
    from ggs_b200 import synth
    g = torch.from_numpy(synth.new_population_np(P, N, H, W, seed=seed)).cuda()
    t_np = synth.synthetic_target_np(H, W, seed)
    return g, torch.from_numpy(t_np).cuda(), torch.from_numpy(synth.importance_mask_np(t_np)).cuda()


def local_group(world, capacity):
    from ggs_b200.peers import PeerGroup
    groups = [PeerGroup(r, world, capacity, device="cuda") for r in range(world)]
    PeerGroup.connect_local(groups)
    return groups


@pytest.mark.parametrize("world,P", [(1, 9), (2, 37), (3, 8), (4, 1024), (8, 5)])
def test_allgather_equals_one_evaluation(ggs, world, P):
    """Every rank's gathered vector == one evaluation of the whole population, bit for bit;
    ragged and empty shards included (8 ranks, 5 candidates)."""
    from ggs_b200.distributed import shard_bounds
    N, H, W = (80, 96, 128) if P < 1000 else (60, 64, 64)
    g, t, m = inputs(P, N, H, W, seed=world)
    full = ggs.fitness(g, t, H, W, 3.0, weight_mask=m)
    groups = local_group(world, P)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for epoch in range(5):                       # five gathers: both buffer halves, reused
        views = []
        for r in reversed(range(world)):         # enqueue order must not matter
            lo, hi = shard_bounds(P, world, r)
            with torch.cuda.stream(streams[r]):
                views.append((r, groups[r].fitness_allgather(g[lo:hi], t, H, W, offset=lo, total=P,
                                                             weight_mask=m)))
        for r, v in views:
            streams[r].synchronize()
            assert torch.equal(v, full), (world, P, epoch, r)
    for grp in groups:
        grp.check()
        grp.close()


def test_sharded_ga_engine_equals_the_single_gpu_engine(ggs):
    """Three ranks in one process: replicated breeding / ranking, sharded evaluation, the select
    kernel waits for the peers' values itself.  Curves, best individual and final population
    equal the unsharded engine's, bit for bit, on every rank."""
    import modules.config as C
    from ggs_b200.engine import GaEngine
    from modules.utils import build_mut_sigma, scale_log_bounds
    P, N, H, W, n_elite, gens = 26, 50, 64, 96, 4, 12
    pop, t, m = inputs(P, N, H, W, seed=3)
    lo, hi = scale_log_bounds(H, W, 3.0, 0.1)
    rows = [build_mut_sigma(g, gens, "cosine", C.MUT_SIGMA_MAX, C.MUT_SIGMA_MIN) for g in range(1, gens + 1)]

    def run(engine, first, second):
        engine.run(rows[:first], 2, 0.5, 0.05, lo, hi)
        engine.run(rows[first:first + second], 2, 0.5, 0.05, lo, hi)

    ref = GaEngine(t, m, H, W, P, N, n_elite, gens)
    ref.start(pop, 77)
    run(ref, 5, 7)
    want = ref.state()
    want_pop, want_fit = ref.population()
    ref.close()

    world = 3
    groups = local_group(world, P)
    streams = [torch.cuda.Stream() for _ in range(world)]
    engines = []
    for r in range(world):
        streams[r].wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(streams[r]):
            e = GaEngine(t, m, H, W, P, N, n_elite, gens)
            e.set_peers(groups[r])
            engines.append(e)
    # start() and state() synchronise their stream: issue every rank's work before reading any
    import threading
    results = [None] * world

    def drive(r):
        with torch.cuda.stream(streams[r]):
            engines[r].start(pop, 77)
            run(engines[r], 3 + r, 9 - r)        # ranks cut their blocks differently
            st = engines[r].state()
            results[r] = (st, engines[r].population())

    threads = [threading.Thread(target=drive, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
        assert not th.is_alive(), "a rank is stuck waiting for its peers"
    for r in range(world):
        st, (p_r, f_r) = results[r]
        assert st["generation"] == gens
        assert np.array_equal(st["curves"], want["curves"]), r
        assert st["best_fitness"] == want["best_fitness"]
        assert torch.equal(st["best_individual"], want["best_individual"])
        assert torch.equal(p_r, want_pop) and torch.equal(f_r, want_fit)
    for e in engines:
        e.close()
    for grp in groups:
        grp.check()
        grp.close()


def test_peers_argument_checks(ggs):
    import ggs_b200
    from ggs_b200.peers import PeerGroup
    with pytest.raises(ggs_b200.GgsError):
        PeerGroup(2, 2, 8, device="cuda")            # rank outside the world
    with pytest.raises(ggs_b200.GgsError):
        PeerGroup(0, 9, 8, device="cuda")            # more than one box
    lone = PeerGroup(0, 2, 8, device="cuda")         # never connected
    g, t, m = inputs(4, 10, 32, 32, seed=1)
    with pytest.raises(ggs_b200.GgsError):
        lone.fitness_allgather(g, t, 32, 32, offset=0, total=4)
    lone.close()
    one = PeerGroup(0, 1, 4, device="cuda")          # a world of one needs no connection
    with pytest.raises(ggs_b200.GgsError):
        one.fitness_allgather(g, t, 32, 32, offset=2, total=4)     # shard sticks out of the vector
    out = one.fitness_allgather(g, t, 32, 32, offset=0, total=4)
    assert torch.equal(out, ggs_b200.fitness(g, t, 32, 32, 3.0))
    one.close()
