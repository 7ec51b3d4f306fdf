"""Host-buffer entry (ggs_ctx_fitness_host) against the device entry (ggs_fitness) on the same genomes,
bit for bit, over batch sizes from one try to a full population; prints which split of the device
entry the host path equals.   gpurun -- python tools/debug_host_path.py"""
import sys, numpy as np, torch
sys.path.insert(0, "genetic-gaussian-splats_b200")
import ggs_b200
from ggs_b200 import synth
dev = torch.device("cuda", 0)
for (side, N, P) in [(128, 100, 32), (128, 100, 16), (128, 100, 64), (256, 500, 8), (256, 500, 24), (128, 100, 128), (256, 1000, 1024)]:
    H = W = side
    t_np = synth.synthetic_target_np(H, W, 0)
    target = torch.from_numpy(t_np).to(dev)
    mask = ggs_b200.importance_mask(target, H, W, edge_scales=(1, 2, 4), w_edge=0.7, w_var=0.3, gamma=0.7, floor=0.15, smooth=3, strength=0.7)
    g_np = synth.new_population_np(P, N, H, W, seed=42)
    g = torch.from_numpy(g_np).to(dev)
    he = ggs_b200.HostEvaluator(t_np, mask.cpu().numpy(), device=0)
    hp = torch.from_numpy(g_np).pin_memory()
    out = torch.empty((P,), dtype=torch.float32).pin_memory()
    for rep in range(3):
        he.fitness(hp, out=out)
        chk = ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask).cpu()
        d = (chk - out).abs()
        bad = torch.nonzero(d > 0).flatten().tolist()
        print(f"{side} {N} {P} rep {rep}: differing {len(bad)} {bad[:12]} max rel {float((d / chk.abs()).max()):.3e}")
    for s in (1, 2, 4, 8):
        chk = ggs_b200.fitness(g, target, H, W, 3.0, weight_mask=mask, split=s).cpu()
        print(f"   split {s}: equal to host path: {torch.equal(chk, out)}  choose_split {ggs_b200.choose_split(P, N, H, W)}")
    he.close()
