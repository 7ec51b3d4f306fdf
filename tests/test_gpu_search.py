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
# the same through peer-to-peer stores: the raster kernel writes into every rank's vector
from ggs_b200.peers import PeerGroup
peers = PeerGroup.from_process_group(capacity=P, device=dev)
ev2 = ShardedEvaluator(t, H, W, weight_mask=m, device=dev, peers=peers)
for _ in range(5):                                   # epochs alternate between two buffer halves
    c = ev2.fitness(g)
    d = ev2.fitness(g[lo:hi].clone(), replicated=False, total=P)
    ok = ok and torch.equal(c, full) and torch.equal(d, full)
peers.check()
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


GA_WORKER = r'''
import os, sys, torch, torch.distributed as dist
os.environ["TQDM_DISABLE"] = "1"
root = sys.argv[1]
sys.path[:0] = [root, os.path.join(root, "genetic-gaussian-splats_b200")]
import modules.config as C
from modules.algorithm import genetic_approx
from ggs_b200 import synth
H, W = 64, 96
target = torch.from_numpy(synth.synthetic_target_np(H, W, 9))
kw = dict(H=H, W=W, device="cuda", pop_size=22, n_splats=40, generations=15, tour_k=C.TOUR_K,
          elite_k=C.ELITE_K, cxpb=C.CXPB, mutpb=C.MUTPB, mut_sigma_max=C.MUT_SIGMA_MAX,
          mut_sigma_min=C.MUT_SIGMA_MIN, schedule=C.SCHEDULE, min_scale_splats=C.MIN_SCALE_SPLATS,
          max_scale_splats=C.MAX_SCALE_SPLATS, k_sigma=C.K_SIGMA, mask_strength=C.MASK_STRENGTH,
          boost_only=C.BOOST_ONLY)
torch.manual_seed(5 + int(os.environ["RANK"]))      # ranks seeded DIFFERENTLY on purpose
best_s, fit_s = genetic_approx(target, **kw)        # sharded over the ranks (torchrun env)
assert dist.is_initialized() and dist.get_world_size() == int(os.environ["WORLD_SIZE"])
# every rank must hold the same result ...
probe = torch.cat([best_s.flatten().double(), torch.tensor([fit_s], dtype=torch.float64)]).cuda()
lo, hi = probe.clone(), probe.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same = bool(torch.equal(lo, hi))
# ... and it must be the single-GPU run of rank 0's seed, bit for bit
os.environ["GGS_B200_NO_SHARD"] = "1"
torch.manual_seed(5)
best_1, fit_1 = genetic_approx(target, **kw)
ok = same and fit_1 == fit_s and torch.equal(best_1, best_s)
flag = torch.tensor([1 if ok else 0]).cuda()
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 3)
'''


def test_two_gpu_sharded_ga_equals_single_gpu_run(cuda_ok, tmp_path):
    """torchrun + the unchanged GA entry point: replicated breeding, sharded evaluation."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "ga_worker.py"
    script.write_text(GA_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29654", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_run_ggs_flow_end_to_end(cuda_ok, tmp_path):
    """The flow of run_ggs.py:31-77 with its outputs: load an image file, choose the work size,
    run the GA with frames and loss curves, rescale the best genome and render it at full size."""
    import modules.config as C
    from PIL import Image
    from modules.algorithm import genetic_approx
    from modules.encode import genome_to_renderer
    from modules.render import _DEV as DEV, render_splats_rgb_triton
    from modules.resize import choose_work_size, scale_genome_pixels_anisotropic
    H_out, W_out = 96, 144
    src = (hidden_target(H_out, W_out, 60).numpy() * 255).astype("uint8")
    img_path = tmp_path / "reference.jpg"
    Image.fromarray(src).save(img_path)
    np_img = np.array(Image.open(img_path).convert("RGB"), dtype=np.float32) / 255.0
    target_img = torch.from_numpy(np_img)
    H, W = choose_work_size(H_out, W_out, max_side=64)
    assert (H, W) == (43, 64)
    frames = tmp_path / "frames"
    frames.mkdir()
    best, fit = genetic_approx(
        target_img, H=H, W=W, device=DEV, pop_size=16, n_splats=48, generations=12, tour_k=C.TOUR_K,
        elite_k=4, cxpb=C.CXPB, mutpb=C.MUTPB, mut_sigma_max=C.MUT_SIGMA_MAX,
        mut_sigma_min=C.MUT_SIGMA_MIN, schedule=C.SCHEDULE, min_scale_splats=C.MIN_SCALE_SPLATS,
        max_scale_splats=C.MAX_SCALE_SPLATS, k_sigma=C.K_SIGMA, mask_strength=C.MASK_STRENGTH,
        boost_only=C.BOOST_ONLY, save_video=True, frame_every=4, video_dir=str(frames), prefix="ga",
        loss_png_path="", loss_csv_path=str(tmp_path / "out" / "ga_loss.csv"), loss_log_y=True)
    assert best.shape == (48, 9) and 0 < fit < 1
    assert sorted(p.name for p in frames.iterdir()) == ["ga_00.png", "ga_04.png", "ga_08.png", "ga_12.png"]
    rows = (tmp_path / "out" / "ga_loss.csv").read_text().strip().splitlines()
    assert rows[0] == "gen,best,mean,median" and len(rows) == 14
    full = scale_genome_pixels_anisotropic(best.to(DEV), sH=H_out / float(H), sW=W_out / float(W))
    final = render_splats_rgb_triton(genome_to_renderer(full).unsqueeze(0), H_out, W_out,
                                     k_sigma=C.K_SIGMA, device=DEV, tile=C.DEFAULT_TILE_SIZE)[0]
    img8 = (final.clamp(0, 1).detach().cpu().numpy() * 255).astype("uint8")
    assert img8.shape == (H_out, W_out, 3)
    Image.fromarray(img8).save(tmp_path / "ga_splats.png")
