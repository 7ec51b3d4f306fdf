"""Pin the CPU oracle (oracle/ggs_oracle.c) against fixtures produced by the reference
itself: tests/golden/make_golden.py (torch CPU + the Triton interpreter, in the build container)
and tests/golden/make_gpu_golden.py (gpu_*.npz: the reference's real CUDA/Triton path on a B200;
images are kept for the first few candidates only).  CPU only."""
import numpy as np

from oracle import oracle

FLOAT_KEYS = oracle.DECODE_FLOAT_KEYS
INT_KEYS = oracle.DECODE_INT_KEYS


def ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def test_encode_matches_reference(golden):
    # encode.py:63-79; transcendental libraries differ (SLEEF in torch CPU, glibc here),
    # so allow a few ulp on the log/sqrt chain and exactness on the pass-through columns.
    chol = oracle.encode(golden["axes"])
    ref = golden["chol"]
    assert chol.shape == ref.shape
    for col in (0, 1, 5, 6, 7, 8):
        assert np.array_equal(chol[..., col], ref[..., col])
    for col in (2, 3):
        np.testing.assert_allclose(chol[..., col], ref[..., col], rtol=2e-6, atol=1e-6)
    # l21 = (cos sin (sx^2 - sy^2)) / l11 cancels; torch CUDA's sin/cos differ from glibc's too
    np.testing.assert_allclose(chol[..., 4], ref[..., 4], rtol=2e-6, atol=5e-6)


def test_decode_aabb_bit_exact(golden):
    # render.py:27-30: the integer AABB is semantics; it must be identical when the decode
    # is fed the reference's own Cholesky genomes.
    H, W, k = int(golden["H"]), int(golden["W"]), float(golden["k_sigma"])
    dec = oracle.decode(golden["chol"], H, W, k)
    for key in INT_KEYS:
        assert np.array_equal(dec[key], golden["dec_" + key]), key
    for key in FLOAT_KEYS:
        # exp: SLEEF (torch CPU), libdevice (torch CUDA), glibc (here); x/255 is x*(1/255) on CUDA
        assert ulp_diff(dec[key], golden["dec_" + key]).max() <= 8, key


def test_decode_after_own_encode_aabb(golden):
    # Same, but through the oracle's own encode (glibc vs SLEEF last-ulp differences can
    # move an AABB edge only if a bound lands within an ulp of an integer).
    H, W, k = int(golden["H"]), int(golden["W"]), float(golden["k_sigma"])
    dec = oracle.decode(oracle.encode(golden["axes"]), H, W, k)
    mism = sum(int((dec[key] != golden["dec_" + key]).sum()) for key in INT_KEYS)
    assert mism == 0


def test_render_matches_reference_kernel(golden):
    # render.py:204-252 through the reference Triton kernel (interpreter): <= 1e-4 abs
    # is the north-star bound; the dense restatement actually agrees to ~1e-6.
    H, W, k = int(golden["H"]), int(golden["W"]), float(golden["k_sigma"])
    img = oracle.render(golden["chol"][:len(golden["images"])], H, W, k)
    err = np.abs(img - golden["images"]).max()
    assert err <= 2e-6, err


def test_fitness_three_modes(golden):
    # fitness.py:16-31, tolerance 1e-5 relative (north star).
    H, W, k = int(golden["H"]), int(golden["W"]), float(golden["k_sigma"])
    a, t, m = golden["axes"], golden["target"], golden["mask"]
    np.testing.assert_allclose(oracle.fitness(a, t, H, W, k), golden["fit_plain"], rtol=1e-5)
    np.testing.assert_allclose(oracle.fitness(a, t, H, W, k, weight_mask=m), golden["fit_mask"],
                               rtol=1e-5)
    np.testing.assert_allclose(oracle.fitness(a, t, H, W, k, weight_mask=m, boost_only=True),
                               golden["fit_boost"], rtol=1e-5)


def test_masked_fitness_keeps_the_denominator_quirk(golden):
    # fitness.py:29-31: denominator sums w over H*W only -> 3x a per-channel weighted mean.
    H, W, k = int(golden["H"]), int(golden["W"]), float(golden["k_sigma"])
    img = golden["images"].astype(np.float64)
    w = golden["mask"].astype(np.float64)
    d2 = (img - golden["target"][None].astype(np.float64)) ** 2
    quirk = (d2 * w[None, :, :, None]).sum(axis=(1, 2, 3)) / (w.sum() + 1e-12)
    got = oracle.fitness(golden["axes"], golden["target"], H, W, k, weight_mask=golden["mask"])
    np.testing.assert_allclose(got[:len(img)], quirk, rtol=1e-5)


def test_score_of_golden_images(golden):
    got = oracle.score(golden["images"], golden["target"])
    np.testing.assert_allclose(got, golden["fit_plain"][:len(got)], rtol=2e-6)


def test_rank_order_matches(golden):
    H, W, k = int(golden["H"]), int(golden["W"]), float(golden["k_sigma"])
    got = oracle.fitness(golden["axes"], golden["target"], H, W, k, weight_mask=golden["mask"])
    assert np.array_equal(np.argsort(got, kind="stable"),
                          np.argsort(golden["fit_mask"], kind="stable"))
