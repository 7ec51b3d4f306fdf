"""World-size-2 gloo tests (CPU) of the population sharding logic.  The per-rank evaluation
is injected (the CPU oracle) so the sharding, the ragged all-gather and the elite exchange are
exercised without a GPU; the GPU equivalent runs in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_partition():
    from ggs_b200.distributed import owner_of, shard_bounds, shard_sizes
    for total in (0, 1, 7, 32, 1024, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    assert owner_of(0, 10, 3) == 0 and owner_of(4, 10, 3) == 1 and owner_of(9, 10, 3) == 2


def _worker(rank, world, port, P, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ggs_b200 import synth
        from ggs_b200.distributed import ShardedEvaluator, shard_bounds
        from oracle import oracle, torch_ref
        oracle.set_threads(1)
        N, H, W = 12, 24, 40
        g = torch.from_numpy(synth.new_population_np(P, N, H, W, seed=3))
        t = synth.synthetic_target_np(H, W, 3)
        m = torch_ref.importance_mask_np(t)

        def evaluate(x):
            return torch.from_numpy(oracle.fitness(x.numpy(), t, H, W, 3.0, weight_mask=m))

        ev = ShardedEvaluator(torch.from_numpy(t), H, W, weight_mask=torch.from_numpy(m),
                              evaluate=evaluate)
        full = evaluate(g)
        got_rep = ev.fitness(g, replicated=True)
        lo, hi = shard_bounds(P, world, rank)
        got_shard = ev.fitness(g[lo:hi], replicated=False, total=P)
        el = ev.elites(got_rep, 3)
        rows = ev.gather_rows(g[lo:hi].clone(), el, P)
        ok = (torch.equal(got_rep, full) and torch.equal(got_shard, full)
              and el == torch.argsort(full, stable=True)[:3].tolist()
              and torch.equal(rows, g[el]))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [8, 7])  # even and ragged split
def test_two_rank_gloo_matches_single_process(P):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + P) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}
