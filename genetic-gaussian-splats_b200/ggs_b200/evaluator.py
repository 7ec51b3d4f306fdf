"""Torch-facing layer over the C ABI: device memory, streams and workspaces only.

PyTorch is plumbing here (allocation, current stream, device guard); every computation
is a kernel of libggs_b200.so.  Entry points mirror the reference functions:

  encode()   genome_to_renderer_batched   modules/encode.py:63-79
  decode()   _preprocess_genome           modules/render.py:9-47
  render()   render_splats_rgb_triton     modules/render.py:204-252
  fitness()  fitness_many                 modules/fitness.py:8-31
  HostEvaluator.fitness()  fitness_population on host buffers   modules/fitness.py:35-48
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import native
from .native import (LAYOUT_AXES_ANGLE, LAYOUT_CHOLESKY, MODE_BOOST, MODE_MASK, MODE_PLAIN,
                     check, lib)

DECODE_FLOAT_KEYS = ("cx", "cy", "sxx", "sxy", "syy", "rc", "gc", "bc", "a")
DECODE_INT_KEYS = ("x0", "x1", "y0", "y1")

_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}
_retired: list = []


def _cuda_device(device) -> torch.device:
    dev = torch.device(device if device is not None else "cuda")
    assert dev.type == "cuda", "This renderer requires a CUDA device."
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _as_f32(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch per (device, stream); the library itself never allocates."""
    key = (dev.index, _stream_ptr(dev))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            # a captured CUDA graph may still point at the old buffer: retire it, never free it
            _retired.append(ws)
            nbytes = max(nbytes, ws.numel() * 3 // 2)   # geometric growth bounds what is retired
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def mode_of(weight_mask, boost_only: bool) -> int:
    if weight_mask is None:
        return MODE_PLAIN
    return MODE_BOOST if boost_only else MODE_MASK


@torch.no_grad()
def encode(axes: torch.Tensor, device=None) -> torch.Tensor:
    """[..., C>=9] axes-angle genomes -> [..., 9] Cholesky layout (encode.py:63-79)."""
    dev = _cuda_device(device if device is not None else axes.device)
    a = _as_f32(axes, dev)
    cols = a.shape[-1]
    assert cols >= 9, "expected at least 9 genome cols"
    rows = a.numel() // cols
    out = torch.empty(a.shape[:-1] + (9,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib().ggs_encode(a.data_ptr(), rows, cols, out.data_ptr(), _stream_ptr(dev)),
              "ggs_encode")
    return out


@torch.no_grad()
def decode(genomes: torch.Tensor, H: int, W: int, k_sigma: float = 3.0,
           layout: int = LAYOUT_CHOLESKY, device=None) -> Dict[str, torch.Tensor]:
    """The reference's 13 per-splat arrays (render.py:9-47), for tests and tools."""
    dev = _cuda_device(device if device is not None else genomes.device)
    g = _as_f32(genomes, dev)
    cols = g.shape[-1]
    rows = g.numel() // cols
    of = torch.empty((9, rows), dtype=torch.float32, device=dev)
    oi = torch.empty((4, rows), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().ggs_decode(g.data_ptr(), layout, rows, cols, int(H), int(W), float(k_sigma),
                               of.data_ptr(), oi.data_ptr(), _stream_ptr(dev)), "ggs_decode")
    shape = g.shape[:-1]
    out = {k: of[i].reshape(shape) for i, k in enumerate(DECODE_FLOAT_KEYS)}
    out.update({k: oi[i].reshape(shape) for i, k in enumerate(DECODE_INT_KEYS)})
    return out


@torch.no_grad()
def render(genomes: torch.Tensor, H: int, W: int, k_sigma: float = 3.0,
           background=(1.0, 1.0, 1.0), layout: int = LAYOUT_CHOLESKY, device=None,
           as_uint8: bool = False) -> torch.Tensor:
    """[B,N,C>=9] genomes -> [B,H,W,3] float32 in [0,1] (render.py:204-252), or uint8
    (image * 255 truncated, utils.py:57) when as_uint8."""
    dev = _cuda_device(device if device is not None else genomes.device)
    g = _as_f32(genomes, dev)
    assert g.ndim == 3 and g.shape[2] >= 9
    B, N, C = g.shape
    img = torch.empty((B, H, W, 3), dtype=torch.uint8 if as_uint8 else torch.float32, device=dev)
    bg = (ctypes.c_float * 3)(*[float(v) for v in background])
    nbytes = lib().ggs_workspace_bytes(B, N, int(H), int(W))
    entry = lib().ggs_render_u8 if as_uint8 else lib().ggs_render
    with torch.cuda.device(dev):
        ws = _workspace(dev, nbytes)
        check(entry(g.data_ptr(), layout, B, N, C, int(H), int(W), float(k_sigma), bg,
                    img.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "ggs_render")
    return img


@torch.no_grad()
def fitness(genomes: torch.Tensor, target: torch.Tensor, H: int, W: int, k_sigma: float = 3.0,
            weight_mask: Optional[torch.Tensor] = None, boost_only: bool = False,
            boost_beta: float = 1.0, layout: int = LAYOUT_AXES_ANGLE, want_images: bool = False,
            device=None, split: int = 0):
    """[B,N,C>=9] genomes -> [B] float32 fitness, fused on device (fitness.py:8-31).

    split: CTAs per (candidate, tile) of the small-batch path, 0 = chosen from B
    (choose_split).  Pass the split of the WHOLE population when it is evaluated in several
    calls or on several GPUs and the bits of a single call are wanted."""
    dev = _cuda_device(device if device is not None else genomes.device)
    g = _as_f32(genomes, dev)
    assert g.ndim == 3 and g.shape[2] >= 9
    B, N, C = g.shape
    t = _as_f32(target, dev)
    assert tuple(t.shape) == (H, W, 3), f"target must be [H,W,3], got {tuple(t.shape)}"
    m = None if weight_mask is None else _as_f32(weight_mask, dev)
    if m is not None:
        assert tuple(m.shape) == (H, W), f"weight_mask must be [H,W], got {tuple(m.shape)}"
    fit = torch.empty((B,), dtype=torch.float32, device=dev)
    img = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev) if want_images else None
    nbytes = lib().ggs_workspace_bytes(B, N, int(H), int(W))
    with torch.cuda.device(dev):
        ws = _workspace(dev, nbytes)
        check(lib().ggs_fitness_ex(g.data_ptr(), layout, B, N, C, int(H), int(W), float(k_sigma),
                                   t.data_ptr(), None if m is None else m.data_ptr(),
                                   mode_of(m, boost_only), float(boost_beta), fit.data_ptr(),
                                   None if img is None else img.data_ptr(), ws.data_ptr(),
                                   ws.numel(), int(split), _stream_ptr(dev)), "ggs_fitness")
    return (fit, img) if want_images else fit


def choose_split(B: int, N: int, H: int, W: int) -> int:
    """The split (1, 2, 4 or 8 CTAs per tile) the library picks for a batch of B candidates."""
    return int(lib().ggs_choose_split(int(B), int(N), int(H), int(W)))


def set_option(name: str, value: int) -> None:
    """Process-wide switch of the library: "pdl", "split", "fuse", "tile_order" (ggs_set_option)."""
    check(lib().ggs_set_option(name.encode(), int(value)), "ggs_set_option")


def tile_order(ntx: int, nty: int):
    """The centre-out order of an ntx x nty tile grid as a list of (tx, ty) (ggs_tile_order; host only)."""
    out = (ctypes.c_int * (2 * ntx * nty))()
    check(lib().ggs_tile_order(int(ntx), int(nty), out), "ggs_tile_order")
    return [(out[2 * r], out[2 * r + 1]) for r in range(ntx * nty)]


def probe_peaks() -> dict:
    """FP32 FMA / FFMA2 / MUFU.EX2 rates measured on the current device (for bench.py)."""
    out = (ctypes.c_float * 5)()
    check(lib().ggs_probe_peaks(out), "ggs_probe_peaks")
    return {"ffma_tflops": out[0], "ffma2_tflops": out[1], "mufu_ex2_gops": out[2],
            "sm_count": int(out[3]), "sm_clock_mhz": out[4]}


SIGMA_ORDER = ("xy", "alog", "blog", "theta", "rgb", "alpha")


@torch.no_grad()
def breed(population: torch.Tensor, fitness_values: torch.Tensor, sigma: dict, *, tour_k: int,
          cxpb: float, mutpb: float, log_scale_lo: float, log_scale_hi: float, seed: int,
          generation: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One GA breeding step on device (ggs_ga_breed): [P,N,C>=9] + [P] -> offspring [P,N,9]."""
    dev = _cuda_device(population.device)
    pop = _as_f32(population, dev)
    fit = _as_f32(fitness_values, dev)
    P, N, C = pop.shape
    assert fit.shape == (P,)
    if out is None:
        out = torch.empty((P, N, 9), dtype=torch.float32, device=dev)
    assert out.shape == (P, N, 9) and out.is_contiguous() and out.data_ptr() != pop.data_ptr()
    sig = (ctypes.c_float * 6)(*[float(sigma[k]) for k in SIGMA_ORDER])
    with torch.cuda.device(dev):
        check(lib().ggs_ga_breed(pop.data_ptr(), fit.data_ptr(), P, N, C, out.data_ptr(),
                                 int(tour_k), float(cxpb), float(mutpb), sig, float(log_scale_lo),
                                 float(log_scale_hi), int(seed) & (2**64 - 1),
                                 int(generation) & 0xffffffff, _stream_ptr(dev)), "ggs_ga_breed")
    return out


def importance_mask(image: torch.Tensor, H: int, W: int, *, edge_scales=(1, 2, 4),
                    w_edge: float = 0.7, w_var: float = 0.3, gamma: float = 0.7,
                    floor: float = 0.15, smooth: int = 0, strength: float = 1.0) -> torch.Tensor:
    """compute_importance_mask (mask.py:29-83) on the device: [H0,W0,3] image -> [H,W] weights
    (ggs_importance_mask).  Same parameters and defaults as the reference."""
    dev = _cuda_device(image.device)
    img = _as_f32(image, dev)
    assert img.ndim == 3 and img.shape[2] == 3, "image must be [H0, W0, 3]"
    H0, W0 = int(img.shape[0]), int(img.shape[1])
    H, W = int(H), int(W)
    is255 = bool(img.max() > 1.5)                      # mask.py:45
    scales = (ctypes.c_int * len(edge_scales))(*[int(s) for s in edge_scales])
    out = torch.empty((H, W), dtype=torch.float32, device=dev)
    need = lib().ggs_mask_workspace_bytes(H, W)
    ws = torch.empty(max(int(need), 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().ggs_importance_mask(img.data_ptr(), H0, W0, H, W, int(is255), scales,
                                        len(edge_scales), float(w_edge), float(w_var), float(gamma),
                                        float(floor), int(smooth or 0), float(strength),
                                        out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev)),
              "ggs_importance_mask")
    ws.record_stream(torch.cuda.current_stream(dev))
    return out


def count_evaluated_pairs(run, device=None) -> dict:
    """Run `run()` (any evaluation) on the instrumented raster kernel and return the number
    of (pixel, splat) pairs it actually evaluated, split by path."""
    dev = _cuda_device(device)
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    check(lib().ggs_stats_target(counters.data_ptr()), "ggs_stats_target")
    try:
        run()
        torch.cuda.synchronize(dev)
    finally:
        check(lib().ggs_stats_target(None), "ggs_stats_target")
    rec, exact = (int(v) * 64 for v in counters.tolist())
    return {"recurrence_pairs": rec, "exact_pairs": exact, "pairs": rec + exact}


def timing_enable(enable: bool = True) -> None:
    check(lib().ggs_timing_enable(1 if enable else 0), "ggs_timing_enable")


def timing_read() -> dict:
    """Summed decode / raster kernel milliseconds since timing_enable (CUDA events)."""
    dec, ras, n = ctypes.c_float(), ctypes.c_float(), ctypes.c_int()
    check(lib().ggs_timing_read(ctypes.byref(dec), ctypes.byref(ras), ctypes.byref(n)),
          "ggs_timing_read")
    return {"decode_ms": dec.value, "raster_ms": ras.value, "evaluations": n.value}


class HostEvaluator:
    """fitness_population on HOST buffers through ggs_ctx_* (fitness.py:35-48).

    The target and mask are uploaded once and stay resident; every fitness() call copies
    the genomes host->device (overlapped with compute), evaluates, and copies B floats back.
    """

    def __init__(self, target, weight_mask=None, device: int = 0):
        t = np.ascontiguousarray(np.asarray(target, dtype=np.float32))
        assert t.ndim == 3 and t.shape[2] == 3
        self.H, self.W = int(t.shape[0]), int(t.shape[1])
        m = None
        if weight_mask is not None:
            m = np.ascontiguousarray(np.asarray(weight_mask, dtype=np.float32))
            assert m.shape == (self.H, self.W)
        self.has_mask = m is not None
        self._ctx = ctypes.c_void_p()
        check(lib().ggs_ctx_create(int(device), ctypes.byref(self._ctx)), "ggs_ctx_create")
        check(lib().ggs_ctx_set_target(self._ctx, t.ctypes.data, None if m is None else m.ctypes.data,
                                       self.H, self.W), "ggs_ctx_set_target")

    def fitness(self, genomes, k_sigma: float = 3.0, boost_only: bool = False,
                boost_beta: float = 1.0, layout: int = LAYOUT_AXES_ANGLE, use_mask: bool = True,
                out=None):
        """genomes: host float32 [B,N,C] (numpy array or CPU torch tensor, ideally pinned)."""
        if isinstance(genomes, torch.Tensor):
            assert genomes.device.type == "cpu" and genomes.dtype == torch.float32
            g = genomes.contiguous()
            B, N, C = g.shape
            gptr = g.data_ptr()
        else:
            g = np.ascontiguousarray(np.asarray(genomes, dtype=np.float32))
            B, N, C = g.shape
            gptr = g.ctypes.data
        if out is None:
            out = np.empty((B,), dtype=np.float32)
        optr = out.data_ptr() if isinstance(out, torch.Tensor) else out.ctypes.data
        mode = MODE_PLAIN if not (use_mask and self.has_mask) else (MODE_BOOST if boost_only else MODE_MASK)
        check(lib().ggs_ctx_fitness_host(self._ctx, gptr, layout, B, N, C, float(k_sigma), mode,
                                         float(boost_beta), optr), "ggs_ctx_fitness_host")
        return out

    def close(self):
        if self._ctx:
            lib().ggs_ctx_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
