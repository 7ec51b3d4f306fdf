"""GPU tests of the callers on either side of the hot path: the batched GA / SA loops through
the drop-in `modules` package, and the sharded evaluator (NCCL when 2+ GPUs are visible)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cuda_ok():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ggs_b200
    ggs_b200.lib()
    return True


def hidden_target(H, W, n, seed=7):
    from modules.encode import genome_to_renderer_batched
    from modules.population import new_population
    from modules.render import render_splats_rgb_triton
    torch.manual_seed(seed)
    g = new_population(1, n, H, W, 3.0, 0.1, device="cuda")
    return render_splats_rgb_triton(genome_to_renderer_batched(g), H, W, k_sigma=3.0, device="cuda")[0].cpu()


def test_ga_improves_fitness(cuda_ok):
    import modules.config as C
    from modules.algorithm import genetic_approx
    H = W = 64
    target = hidden_target(H, W, 40)
    torch.manual_seed(0)
    kw = dict(pop_size=32, n_splats=40, tour_k=2, elite_k=4, cxpb=0.05, mutpb=0.05,
              mut_sigma_max=C.MUT_SIGMA_MAX, mut_sigma_min=C.MUT_SIGMA_MIN, schedule="cosine",
              min_scale_splats=3.0, max_scale_splats=0.1, k_sigma=3.0, mask_strength=0.7,
              boost_only=False)
    _, f0 = genetic_approx(target, H, W, "cuda", generations=0, **kw)
    torch.manual_seed(0)
    best, f1 = genetic_approx(target, H, W, "cuda", generations=60, **kw)
    assert best.shape == (40, 9) and best.device.type == "cpu"
    assert f1 < f0 * 0.9, (f0, f1)          # elitism: monotone best, and 60 generations help


def test_sa_runs_batched_and_sequential(cuda_ok):
    import modules.config as C
    from modules.annealing import simulated_annealing
    H, W = 48, 64
    target = hidden_target(H, W, 30)
    kw = dict(n_splats=30, mutpb=0.05, mut_sigma_max=C.MUT_SIGMA_MAX, mut_sigma_min=C.MUT_SIGMA_MIN,
              sigma_schedule="cosine", min_scale_splats=3.0, max_scale_splats=0.1, k_sigma=3.0,
              mask_strength=0.7, boost_only=False, temp0=1e-3, temp_schedule="cosine",
              tries_per_iter=8)
    for batched in (True, False):
        torch.manual_seed(1)
        _, e0 = simulated_annealing(target, H, W, "cuda", iterations=0, batch_neighbors=batched, **kw)
        torch.manual_seed(1)
        best, e1 = simulated_annealing(target, H, W, "cuda", iterations=40, batch_neighbors=batched, **kw)
        assert best.shape == (30, 9) and e1 <= e0


def test_sharded_evaluator_single_rank_equals_direct(cuda_ok):
    import ggs_b200
    from ggs_b200 import synth
    from ggs_b200.distributed import ShardedEvaluator
    B, N, H, W = 40, 64, 96, 64
    g = torch.from_numpy(synth.new_population_np(B, N, H, W, seed=2)).cuda()
    t_np = synth.synthetic_target_np(H, W, 2)
    t, m = torch.from_numpy(t_np).cuda(), torch.from_numpy(synth.importance_mask_np(t_np)).cuda()
    ev = ShardedEvaluator(t, H, W, weight_mask=m, device="cuda")
    assert torch.equal(ev.fitness(g), ggs_b200.fitness(g, t, H, W, 3.0, weight_mask=m))


NCCL_WORKER = r'''
import os, sys, torch, torch.distributed as dist
root = sys.argv[1]
sys.path[:0] = [root, os.path.join(root, "genetic-gaussian-splats_b200")]
import ggs_b200
from ggs_b200 import synth
from ggs_b200.distributed import ShardedEvaluator, shard_bounds
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
P, N, H, W = 37, 80, 96, 128
g = torch.from_numpy(synth.new_population_np(P, N, H, W, seed=4)).to(dev)
t_np = synth.synthetic_target_np(H, W, 4)
t = torch.from_numpy(t_np).to(dev); m = torch.from_numpy(synth.importance_mask_np(t_np)).to(dev)
ev = ShardedEvaluator(t, H, W, weight_mask=m, device=dev)
full = ggs_b200.fitness(g, t, H, W, 3.0, weight_mask=m)
lo, hi = shard_bounds(P, world, rank)
a = ev.fitness(g)                                   # replicated population, ragged shards
b = ev.fitness(g[lo:hi].clone(), replicated=False, total=P)
el = ev.elites(a, 5)
rows = ev.gather_rows(g[lo:hi].clone(), el, P)
ok = torch.equal(a, full) and torch.equal(b, full) and torch.equal(rows, g[el])
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 3)
'''


def test_two_gpu_nccl_gather_is_bit_identical(cuda_ok, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "nccl_worker.py"
    script.write_text(NCCL_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29653", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
