"""The small-batch (latency) path of the raster: a thread-block cluster of `split` CTAs per
(candidate, tile), each compositing one segment of the genome, partial states folded through
distributed shared memory; optionally with the decode fused into the same launch.  Same parity
bar as the throughput path (image <= 1e-4, fitness <= 1e-5 relative against the oracle), plus:
fused and two-kernel variants of one split agree BIT FOR BIT, a given split is independent of
how the batch is cut into calls, and the automatic choice is the documented function of B."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4
FIT_RTOL = 1e-5


@pytest.fixture(scope="module")
def ggs():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ggs_b200
    ggs_b200.lib()
    yield ggs_b200
    ggs_b200.set_option("fuse", 0)
    ggs_b200.set_option("split", 0)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


CASES = [
    # (B, N, H, W, seed, late)
    (8, 100, 128, 128, 42, False),     # BASELINE config 1 shape
    (2, 500, 256, 256, 43, False),     # config 2 shape
    (1, 500, 256, 256, 44, True),      # the reference's sequential SA try: B = 1
    (3, 37, 90, 70, 45, False),        # ragged image, segments of 4-5 splats at split 8
    (1, 1500, 64, 96, 46, False),      # segments longer than one scan round; no fusion at split <= 2
    (2, 5, 40, 40, 47, False),         # fewer splats than CTAs in the cluster: empty segments
    (1, 4100, 300, 200, 48, True),     # segments of 513+ splats at split 8: several flushes per CTA
]


@pytest.mark.parametrize("split", [1, 2, 4, 8])
@pytest.mark.parametrize("B,N,H,W,seed,late", CASES)
def test_split_and_fused_against_oracle(ggs, B, N, H, W, seed, late, split):
    from ggs_b200 import synth
    g = (synth.late_population_np if late else synth.new_population_np)(B, N, H, W, seed)
    t = synth.synthetic_target_np(H, W, seed)
    m = np.random.default_rng(seed).uniform(0.2, 1.0, size=(H, W)).astype(np.float32)
    fit_cpu, img_cpu = oracle.fitness(g, t, H, W, 3.0, weight_mask=m, return_images=True)
    got = {}
    for fuse in (0, 1):
        ggs.set_option("fuse", fuse)
        fit, img = ggs.fitness(cuda(g), cuda(t), H, W, 3.0, weight_mask=cuda(m), want_images=True,
                               split=split)
        got[fuse] = (fit.cpu().numpy(), img.cpu().numpy())
        assert np.abs(got[fuse][1] - img_cpu).max() <= IMG_TOL, (fuse, split)
        np.testing.assert_allclose(got[fuse][0], fit_cpu, rtol=FIT_RTOL)
        # the three modes share the kernel; check the other two reductions as well
        for kw in ({}, {"weight_mask": m, "boost_only": True}):
            f_cpu = oracle.fitness(g, t, H, W, 3.0, **kw)
            kw_gpu = {k2: (cuda(v) if isinstance(v, np.ndarray) else v) for k2, v in kw.items()}
            f = ggs.fitness(cuda(g), cuda(t), H, W, 3.0, split=split, **kw_gpu).cpu().numpy()
            np.testing.assert_allclose(f, f_cpu, rtol=FIT_RTOL)
    ggs.set_option("fuse", 0)
    # fused decode = decode kernel + raster: the same records, hence the same image bit for bit;
    # the same fitness bits too, except at split 1, where the two-kernel path is the throughput
    # kernel and sums a tile's squared errors in another order
    assert np.array_equal(got[0][1], got[1][1])
    if split > 1:
        assert np.array_equal(got[0][0], got[1][0])
    else:
        np.testing.assert_allclose(got[0][0], got[1][0], rtol=1e-6)


def test_a_given_split_is_deterministic_and_independent_of_the_call_boundaries(ggs):
    from ggs_b200 import synth
    B, N, H, W = 13, 240, 128, 96
    g = cuda(synth.new_population_np(B, N, H, W, seed=5))
    t = cuda(synth.synthetic_target_np(H, W, 5))
    for split in (1, 2, 4, 8):
        for fuse in (0, 1):
            ggs.set_option("fuse", fuse)
            f1 = ggs.fitness(g, t, H, W, 3.0, split=split)
            assert torch.equal(f1, ggs.fitness(g, t, H, W, 3.0, split=split))
            parts = torch.cat([ggs.fitness(g[:3], t, H, W, 3.0, split=split),
                               ggs.fitness(g[3:], t, H, W, 3.0, split=split)])
            assert torch.equal(f1, parts), (split, fuse)
    ggs.set_option("fuse", 0)
    # different splits agree to rounding, not bit for bit (the fold re-associates the blend)
    a, b = ggs.fitness(g, t, H, W, 3.0, split=1), ggs.fitness(g, t, H, W, 3.0, split=8)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-6)


def test_automatic_split_policy(ggs):
    # 8 within half a wave; 2 / 4 up to 7 CTAs per SM while the unsplit grid has fewer than 3 per SM;
    # segments of at least 16 splats; never on deep genomes
    import ggs_b200
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    slots = 8 * sms
    assert ggs.choose_split(1024, 1000, 256, 256) == 1          # config 3: throughput path
    assert ggs.choose_split(1, 500, 256, 256) == 8              # one SA try: 64 tiles x 8
    assert ggs.choose_split(2, 500, 256, 256) == 4              # 8 would be 1,024 CTAs folding 8 states each
    assert ggs.choose_split(3, 500, 256, 256) == 4              # 768 CTAs
    assert ggs.choose_split(6, 500, 256, 256) == 2              # 768 CTAs
    assert ggs.choose_split(32, 100, 128, 128) == 1             # config 1: 512 CTAs already
    assert ggs.choose_split(8, 500, 256, 256) == 1              # config 2, 8 neighbours: likewise
    assert ggs.choose_split(1, 4000, 512, 512) == 1             # deep genome: keep the saturation stop
    for B, N, side in ((32, 100, 128), (8, 500, 256), (4, 20, 64), (3, 1000, 512), (2, 500, 256), (12, 100, 128),
                       (1, 1000, 512), (5, 64, 96)):
        tiles = ((side + 31) // 32) ** 2
        k = ggs.choose_split(B, N, side, side)
        assert k in (1, 2, 4, 8)
        assert k == 1 or -(-N // k) >= 16
        assert k == 1 or B * tiles * k <= slots - slots // 8
        assert k != 8 or B * tiles * 8 <= slots // 2
    # the default entry uses it: same bits as the explicit call
    from ggs_b200 import synth
    g = cuda(synth.new_population_np(4, 300, 128, 128, seed=9))
    t = cuda(synth.synthetic_target_np(128, 128, 9))
    k = ggs.choose_split(4, 300, 128, 128)
    assert k > 1
    assert torch.equal(ggs.fitness(g, t, 128, 128, 3.0), ggs.fitness(g, t, 128, 128, 3.0, split=k))
    # an override for experiments
    ggs.set_option("split", 2)
    assert ggs.choose_split(1024, 1000, 256, 256) == 2
    ggs.set_option("split", 0)
    with pytest.raises(ggs_b200.GgsError):
        ggs.set_option("split", 3)
    with pytest.raises(ggs_b200.GgsError):
        ggs.fitness(g, t, 128, 128, 3.0, split=5)


def test_cta_order_does_not_change_a_bit(ggs):
    """Grids of more than one CTA per SM and at most sixteen waves run tile-major, tiles centre-out
    (the split kernel too); partial sums are stored and added by tile index, so images and fitness
    are those of the candidate-major order.  Shapes: 1-14 waves, less than a wave (BASELINE configs
    1 and 2), cluster splits 2 / 4 / 8, ragged image sizes."""
    from ggs_b200 import synth
    for (B, N, H, W, split) in ((30, 300, 256, 256, 0), (24, 512, 256, 256, 0), (70, 90, 200, 136, 0),
                                (5, 400, 512, 512, 0), (32, 100, 128, 128, 0), (8, 500, 256, 256, 0),
                                (3, 500, 256, 256, 0), (1, 500, 256, 256, 8), (2, 300, 200, 136, 4),
                                (6, 200, 256, 256, 2), (16, 100, 128, 128, 0), (130, 120, 256, 256, 0),
                                (260, 60, 256, 256, 0)):
        g = cuda(synth.new_population_np(B, N, H, W, seed=B))
        t = cuda(synth.synthetic_target_np(H, W, B))
        m = cuda(np.random.default_rng(B).uniform(0.2, 1.0, size=(H, W)).astype(np.float32))
        out = {}
        for order in (1, 0):
            ggs.set_option("tile_order", order)
            out[order] = ggs.fitness(g, t, H, W, 3.0, weight_mask=m, want_images=True, split=split)
        ggs.set_option("tile_order", 1)
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]), (B, N, H, W)
        f_cpu = oracle.fitness(g.cpu().numpy(), t.cpu().numpy(), H, W, 3.0, weight_mask=m.cpu().numpy())
        np.testing.assert_allclose(out[1][0].cpu().numpy(), f_cpu, rtol=FIT_RTOL)


def test_host_path_and_render_entries_on_the_small_batch_path(ggs):
    from ggs_b200 import synth
    B, N, H, W = 3, 200, 128, 128
    g = synth.new_population_np(B, N, H, W, seed=12)
    t = synth.synthetic_target_np(H, W, 12)
    assert ggs.choose_split(B, N, H, W) > 1
    dev = ggs.fitness(cuda(g), cuda(t), H, W, 3.0).cpu().numpy()
    he = ggs.HostEvaluator(t, None)
    assert np.array_equal(he.fitness(g), dev)          # slices use the whole batch's configuration
    assert np.array_equal(he.fitness(g), dev)          # and leave the workspace reusable
    he.close()
    chol = oracle.encode(g)
    ref = oracle.render(chol, H, W, 3.0, background=(0.2, 0.4, 0.6))
    img = ggs.render(cuda(chol), H, W, 3.0, background=(0.2, 0.4, 0.6)).cpu().numpy()
    assert np.abs(img - ref).max() <= IMG_TOL
    u8 = ggs.render(cuda(chol), H, W, 3.0, background=(0.2, 0.4, 0.6), as_uint8=True).cpu().numpy()
    assert np.array_equal(u8, (img * 255.0).astype("uint8"))


@pytest.mark.parametrize("B,N,H,W", [(32, 100, 128, 128), (64, 100, 128, 128), (8, 500, 256, 256),
                                     (16, 100, 128, 128), (24, 512, 256, 256), (1, 500, 256, 256)])
def test_host_path_equals_device_path_on_grids_of_at_most_a_few_waves(ggs, B, N, H, W):
    # the sliced host-buffer entry picks split, decode fusion and CTA order from the WHOLE batch:
    # same bits as one ggs_fitness call (BASELINE configs 1 and 2, the default GA's children, one try)
    from ggs_b200 import synth
    g = synth.new_population_np(B, N, H, W, seed=5)
    t = synth.synthetic_target_np(H, W, 5)
    m = synth.importance_mask_np(t)
    dev = ggs.fitness(cuda(g), cuda(t), H, W, 3.0, weight_mask=cuda(m)).cpu().numpy()
    he = ggs.HostEvaluator(t, m)
    assert np.array_equal(he.fitness(g), dev)
    assert np.array_equal(he.fitness(g), dev)
    he.close()


def test_small_batch_path_in_a_cuda_graph(ggs):
    # cluster launches and the counter memset of the fused path are capturable
    from ggs_b200 import synth
    B, N, H, W = 2, 300, 128, 128
    g = cuda(synth.new_population_np(B, N, H, W, seed=33))
    g2 = cuda(synth.new_population_np(B, N, H, W, seed=34))
    t = cuda(synth.synthetic_target_np(H, W, 33))
    eager1, eager2 = ggs.fitness(g, t, H, W, 3.0), ggs.fitness(g2, t, H, W, 3.0)
    static_g = g.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ggs.fitness(static_g, t, H, W, 3.0)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        out = ggs.fitness(static_g, t, H, W, 3.0)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager1)
    static_g.copy_(g2)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager2)
