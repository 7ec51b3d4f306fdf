#!/usr/bin/env python
"""CPU-only model of how evenly a sub-wave raster grid loads the SMs under different CTA orders.
Tile weights = lanes x rows the composite would blend, from the oracle's AABBs of the synthetic
populations; the deal = CTA i -> SM position i mod 148 (what tools/probe_cta_dealing.cu shows for
full rounds).  Prints the busiest SM's load over the average for: candidate-major, the library's
centre-out order, centre-out with odd rounds reversed, and two LPT plans (with the true per-candidate
weights: the bound of any order; with per-tile mean weights: the bound of any STATIC order)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "genetic-gaussian-splats_b200")]
from oracle import oracle
from ggs_b200 import synth
import ggs_b200

def tile_weights(side, N, B, seed=1):
    """per (candidate, ty, tx): pixel-splat pairs inside the tile (AABB overlap area) -- the composite's work"""
    H = W = side
    g = synth.new_population_np(B, N, H, W, seed=seed)
    d = oracle.decode(oracle.encode(g), H, W, 3.0)
    nt = (side + 31) // 32
    w = np.zeros((B, nt, nt))
    for ty in range(nt):
        for tx in range(nt):
            X0, Y0 = tx * 32, ty * 32
            ox = np.clip(np.minimum(d["x1"], X0 + 31) - np.maximum(d["x0"], X0) + 1, 0, None)
            oy = np.clip(np.minimum(d["y1"], Y0 + 31) - np.maximum(d["y0"], Y0) + 1, 0, None)
            hit = (ox > 0) & (oy > 0)
            # work: rows handled in bands of 8 -> count lanes(32) x rows overlapped, plus a per-entry overhead
            w[:, ty, tx] = (hit * (32 * oy + 40)).sum(axis=1)
    return w, nt

def loads(order, w, nsm=148, slots=8):
    """order: list of (b, ty, tx) by CTA index; the scheduler deals CTA i to SM position i mod nsm
    (full rounds; the last partial round goes to the first positions)."""
    load = np.zeros(nsm)
    for i, (b, ty, tx) in enumerate(order):
        load[i % nsm] += w[b, ty, tx]
    return load

def report(name, order, w):
    l = loads(order, w)
    print(f"   {name:34s} max/avg {l.max() / l.mean():.3f}   min/avg {l.min() / l.mean():.3f}")
    return l.max() / l.mean()

for (side, N, B) in ((256, 500, 8), (128, 100, 32), (256, 512, 12), (256, 1000, 8), (256, 500, 16)):
    w, nt = tile_weights(side, N, B)
    ctas = B * nt * nt
    print(f"{side}x{side}, {N} splats, {B} candidates: {ctas} CTAs, tile weight min/mean/max {w.min():.0f}/{w.mean():.0f}/{w.max():.0f}")
    cand_major = [(b, ty, tx) for b in range(B) for ty in range(nt) for tx in range(nt)]
    co = ggs_b200.tile_order(nt, nt)
    centre_out = [(b, ty, tx) for (tx, ty) in co for b in range(B)]
    report("candidate-major", cand_major, w)
    report("centre-out (current)", centre_out, w)
    # reverse odd full rounds
    o = list(centre_out)
    for r in range(len(o) // 148):
        if r & 1:
            o[r * 148:(r + 1) * 148] = o[r * 148:(r + 1) * 148][::-1]
    report("centre-out, odd rounds reversed", o, w)
    # oracle LPT with true weights given the slot structure (upper bound of what an order can do)
    items = sorted(cand_major, key=lambda t: -w[t])
    nsm = 148
    cap = np.array([len(range(p, ctas, nsm)) for p in range(nsm)])
    load = np.zeros(nsm); used = np.zeros(nsm, int)
    for it in items:
        free = np.where(used < cap)[0]
        p = free[np.argmin(load[free])]
        load[p] += w[it]; used[p] += 1
    print(f"   {'LPT with the true weights':34s} max/avg {load.max() / load.mean():.3f}")
    # proxy LPT: weights from the ring index only (what the library could know without the genome)
    ring = lambda tx, ty: min(tx, ty, nt - 1 - tx, nt - 1 - ty)
    proxy = {}
    for (b, ty, tx) in cand_major:
        proxy[(b, ty, tx)] = 1.0
    # mean weight per (ty,tx) over candidates as "known" proxy (best case for a static plan)
    mean_w = w.mean(axis=0)
    items = sorted(cand_major, key=lambda t: -mean_w[t[1], t[2]])
    load = np.zeros(nsm); used = np.zeros(nsm, int); pl = np.zeros(nsm)
    for it in items:
        free = np.where(used < cap)[0]
        p = free[np.argmin(pl[free])]
        pl[p] += mean_w[it[1], it[2]]; load[p] += w[it]; used[p] += 1
    print(f"   {'LPT with per-tile mean weights':34s} max/avg {load.max() / load.mean():.3f}")
