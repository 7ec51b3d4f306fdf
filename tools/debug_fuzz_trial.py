"""Reproduce one trial of tests/test_gpu_parity.py::test_randomised_shapes_against_oracle and
report where the GPU image departs from the oracle.  usage: debug_fuzz_trial.py TRIAL"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "genetic-gaussian-splats_b200"), ROOT]
import numpy as np, torch
from ggs_b200 import synth, evaluator as ggs
from oracle import oracle

want = int(sys.argv[1])
rng = np.random.default_rng(2024)
for trial in range(want + 1):
    H, W = int(rng.integers(1, 180)), int(rng.integers(1, 180))
    N, B = int(rng.integers(0, 260)), int(rng.integers(1, 5))
    k = float(rng.choice([1.0, 2.0, 3.0, 4.5]))
    g = synth.new_population_np(B, max(N, 1), H, W, seed=100 + trial)[:, :N]
    if N and trial % 3 == 0:
        g[0, 0, 2:4] = rng.uniform(-6.0, 6.0, size=2)
        g[0, N // 2, 8] = 0.0
        g[-1, -1, 0:2] = rng.uniform(-0.5, 1.5, size=2)
    t = synth.synthetic_target_np(H, W, trial)
    m = rng.uniform(0.2, 1.3, size=(H, W)).astype(np.float32)
    bg = tuple(float(v) for v in rng.uniform(0, 1, size=3))
print("trial", trial, H, W, N, B, k)
chol = oracle.encode(g)
dev = torch.device("cuda:0")
def both(ch):
    ref = oracle.render(ch, H, W, k, bg)
    out = ggs.render(torch.from_numpy(ch).to(dev), H, W, k_sigma=k, background=bg).cpu().numpy()
    return ref, out
ref, out = both(chol)
d = np.abs(out - ref).max(axis=-1)
print("max diff", d.max(), "candidates", d.reshape(B, -1).max(axis=1))
b, y, x = np.unravel_index(np.argmax(d), d.shape)
print("worst pixel b,y,x", b, y, x, "ref", ref[b, y, x], "gpu", out[b, y, x])
thr = 1e-4 if d.max() > 1e-4 else 0.5 * float(d.max())
ys, xs = np.nonzero(d[b] > thr)
print(f"pixels over {thr:.1e}:", len(ys), "y", ys.min(), ys.max(), "x", xs.min(), xs.max())
# which splat: render each splat of candidate b alone
dec = oracle.decode(chol, H, W, k)
worst = []
for i in range(N):
    r1, o1 = both(np.ascontiguousarray(chol[b:b + 1, i:i + 1]))
    e = np.abs(r1 - o1).max()
    if e > 1e-5:
        worst.append((float(e), i))
worst.sort(reverse=True)
print("single-splat errors:", worst[:5])
for e, i in worst[:3]:
    print(i, {kk: dec[kk][b, i] for kk in dec}, "gene", g[b, i])
    dg = ggs.decode(torch.from_numpy(chol[b:b+1, i:i+1]).to(dev), H, W, k_sigma=k)
    print("  gpu decode:", {kk: (v.cpu().numpy().ravel()[0]) for kk, v in dg.items()} if isinstance(dg, dict) else dg)
