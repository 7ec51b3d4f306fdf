"""CPU-side tests: the C-ABI library loads and exports every declared symbol (no compute
calls without a GPU), the host logic of the drop-in `modules` package, and the synthetic
input generators."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ggs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ggs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ggs_b200
    from ggs_b200 import native
    L = ggs_b200.lib()
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ggs_b200.h but not exported"
        assert n in native.SIGNATURES, f"{n} has no ctypes signature in native.py"
    assert L.ggs_abi_version() == native.ABI_VERSION


def test_ctypes_signatures_match_the_header_prototypes():
    """Argument counts and the obvious argument kinds of native.SIGNATURES follow the prototypes
    in include/ggs_b200.h (a drifted binding would corrupt the call without any GPU to notice)."""
    from ggs_b200 import native
    text = open(os.path.join(ROOT, "include", "ggs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(ggs_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) >= 30
    ctype_of = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
                "double": ctypes.c_double, "size_t": ctypes.c_size_t, "uint64_t": ctypes.c_uint64,
                "uint32_t": ctypes.c_uint32}
    for ret, name, params in protos:
        res, args = native.SIGNATURES[name]
        plist = [] if params.strip() in ("", "void") else [q.strip() for q in params.split(",")]
        assert len(plist) == len(args), f"{name}: header has {len(plist)} parameters, binding {len(args)}"
        for q, a in zip(plist, args):
            if "*" in q:    # any pointer: void_p or a typed POINTER
                assert a is ctypes.c_void_p or a is ctypes.c_char_p or hasattr(a, "_type_") and \
                    issubclass(a, ctypes._Pointer), f"{name}: `{q}` bound as {a}"
            else:
                want = q.replace("const ", "").split()[0]
                assert a is ctype_of[want], f"{name}: `{q}` bound as {a}"
        ret = ret.replace("const", "").strip()
        if "*" in ret:
            assert res in (ctypes.c_char_p, ctypes.c_void_p)
        elif ret == "void":
            assert res is None
        else:
            assert res is ctype_of[ret], f"{name}: returns {ret}, bound as {res}"


def test_workspace_size_is_monotone_and_nonzero():
    import ggs_b200
    L = ggs_b200.lib()
    a = L.ggs_workspace_bytes(1, 1, 8, 8)
    b = L.ggs_workspace_bytes(1024, 1000, 256, 256)
    c = L.ggs_workspace_bytes(8192, 4000, 512, 512)
    assert 0 < a < b < c
    # records (48 B) + packed AABBs (8 B) dominate: about 56 B per splat, plus 64 B of partial
    # sums per (candidate, tile) for the up-to-8-way split of the latency path
    assert 56 * 1024 * 1000 <= b <= 56 * 1024 * 1000 + 1024 * 64 * 64 + 8192
    assert L.ggs_workspace_bytes(-1, 1, 8, 8) == 0


def test_missing_library_fails_loudly(monkeypatch):
    from ggs_b200 import native
    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setattr(native, "LIB_PATH", "/nonexistent/libggs_b200.so")
    with pytest.raises(native.GgsError, match="no CPU fallback"):
        native.lib()


def test_render_entry_keeps_the_reference_asserts():
    from modules.render import _DEV, render_splats_rgb_triton
    assert _DEV == "cuda"
    g = torch.zeros((1, 2, 9))
    with pytest.raises(AssertionError, match="CUDA device"):
        render_splats_rgb_triton(g, 8, 8, device="cpu")
    with pytest.raises(AssertionError, match="genomes must be"):
        render_splats_rgb_triton(torch.zeros(9), 8, 8)
    with pytest.raises(AssertionError, match="at least 9"):
        render_splats_rgb_triton(torch.zeros((1, 2, 8)), 8, 8)


def test_entry_signatures_match_the_reference():
    import inspect
    from modules.fitness import fitness_many, fitness_population
    from modules.render import render_splats_rgb_triton
    assert list(inspect.signature(render_splats_rgb_triton).parameters) == [
        "genomes", "H", "W", "k_sigma", "device", "background", "tile", "num_warps",
        "num_stages", "use_fp16_canvas"]
    assert list(inspect.signature(fitness_many).parameters) == [
        "pop_batch", "target", "H", "W", "k_sigma", "device", "tile", "weight_mask",
        "boost_only", "boost_beta"]
    assert list(inspect.signature(fitness_population).parameters) == [
        "population", "target", "H", "W", "k_sigma", "device", "tile", "chunk", "weight_mask",
        "boost_only"]
    sig = inspect.signature(render_splats_rgb_triton)
    assert sig.parameters["tile"].default == 64 and sig.parameters["k_sigma"].default == 3.0
    assert inspect.signature(fitness_many).parameters["tile"].default == 32


def test_mode_selection():
    from ggs_b200 import MODE_BOOST, MODE_MASK, MODE_PLAIN, mode_of
    assert mode_of(None, False) == MODE_PLAIN and mode_of(None, True) == MODE_PLAIN
    assert mode_of(object(), False) == MODE_MASK and mode_of(object(), True) == MODE_BOOST


def test_synthetic_population_follows_new_population():
    from ggs_b200 import synth
    g = synth.new_population_np(16, 500, 256, 192, seed=42)
    assert g.shape == (16, 500, 9) and g.dtype == np.float32
    assert g[..., 0:2].min() >= 0 and g[..., 0:2].max() <= 1
    sig = np.exp(g[..., 2:4])
    assert sig.min() >= 3.0 - 1e-4 and sig.max() <= 25.6 + 1e-3      # [3 px, 0.1*max(H,W)]
    assert abs(sig[..., 0].mean() - (3 + 0.4 * 22.6)) < 0.5          # Beta mean 0.4
    assert abs(sig[..., 1].mean() - (3 + 0.6 * 22.6)) < 0.5          # Beta mean 0.6
    assert np.abs(g[..., 4]).max() <= np.pi + 1e-6
    assert g[..., 8].min() >= 180 and g[..., 5:9].max() <= 255
    assert np.array_equal(g, synth.new_population_np(16, 500, 256, 192, seed=42))


def test_work_statistics_match_the_survey():
    # 128x128, 100 splats -> 2.32e5 in-AABB pairs per candidate
    from ggs_b200 import synth
    from oracle import oracle
    g = synth.new_population_np(32, 100, 128, 128, seed=42)
    d = oracle.decode(oracle.encode(g), 128, 128)
    pairs = synth.count_pairs(d["x0"], d["x1"], d["y0"], d["y1"]) / 32
    assert 2.0e5 < pairs < 2.6e5


def test_mask_matches_reference_formula_on_fixture(golden):
    # the goldens carry the reference's own mask for their synthetic target
    from oracle import torch_ref
    m = torch_ref.importance_mask_np(golden["target"], strength=0.7)
    np.testing.assert_allclose(m, golden["mask"], atol=1e-6)


def test_run_scripts_can_import_everything_they_need():
    # run_ggs.py:8-12 and run_sags.py:8-12 import these names and call the loops by keyword
    import inspect
    from modules.algorithm import genetic_approx
    from modules.annealing import simulated_annealing
    from modules.config import (BOOST_ONLY, CXPB, DEFAULT_TILE_SIZE, ELITE_K, FRAME_EVERY, GENERATIONS,  # noqa: F401
                                INPUT_DIR, K_SIGMA, LOSS_LOG_Y, MASK_STRENGTH, MAX_SCALE_SPLATS,
                                MIN_SCALE_SPLATS, MUT_SIGMA_MAX, MUT_SIGMA_MIN, MUTPB, N_SPLATS,
                                OUTPUT_DIR, POP_SIZE, REF_IMG, SA_SCHEDULE, SA_T0, SA_TRIES_PER_ITER,
                                SAVE_LOSS_CURVE, SAVE_VIDEO, SCHEDULE, SEED, TOUR_K, WORK_MAX_SIDE)
    from modules.encode import genome_to_renderer  # noqa: F401
    from modules.render import _DEV, render_splats_rgb_triton  # noqa: F401
    from modules.resize import choose_work_size, scale_genome_pixels_anisotropic  # noqa: F401
    ga = list(inspect.signature(genetic_approx).parameters)
    assert ga == ["target_img_uint8", "H", "W", "device", "pop_size", "n_splats", "generations",
                  "tour_k", "elite_k", "cxpb", "mutpb", "mut_sigma_max", "mut_sigma_min", "schedule",
                  "min_scale_splats", "max_scale_splats", "k_sigma", "mask_strength", "boost_only",
                  "save_video", "frame_every", "video_dir", "prefix", "loss_png_path",
                  "loss_csv_path", "loss_log_y"]
    sa = list(inspect.signature(simulated_annealing).parameters)
    assert sa[:26] == ["target_img_uint8", "H", "W", "device", "n_splats", "mutpb", "mut_sigma_max",
                       "mut_sigma_min", "sigma_schedule", "min_scale_splats", "max_scale_splats",
                       "k_sigma", "mask_strength", "boost_only", "iterations", "temp0",
                       "temp_schedule", "tries_per_iter", "save_video", "frame_every", "video_dir",
                       "prefix", "loss_png_path", "loss_csv_path", "loss_log_y", "batch_neighbors"]


def test_centre_out_tile_order_is_a_permutation_from_the_centre_outwards():
    """ggs_tile_order (host only): the CTA order of grids between two CTAs per SM and sixteen waves.
    Every tile exactly once; the distance from the image border never increases along the order;
    the four image corners come last."""
    import ggs_b200
    for ntx in range(1, 34):
        for nty in (1, 2, 3, 4, 5, 8, 13, 16, 31, 32):
            order = ggs_b200.tile_order(ntx, nty)
            assert sorted(order) == [(x, y) for x in range(ntx) for y in range(nty)], (ntx, nty)
            ring = [min(x, y, ntx - 1 - x, nty - 1 - y) for x, y in order]
            assert all(a >= b for a, b in zip(ring, ring[1:])), (ntx, nty)
            if ntx >= 3 and nty >= 3:
                assert set(order[-4:]) == {(0, 0), (ntx - 1, 0), (0, nty - 1), (ntx - 1, nty - 1)}
    assert ggs_b200.tile_order(8, 8)[:4] == [(3, 3), (4, 3), (3, 4), (4, 4)]
