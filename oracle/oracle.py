"""ctypes front-end of the CPU oracle (oracle/ggs_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of ggs_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs and by
nothing under genetic-gaussian-splats_b200/.

Reference functions restated (citations are into /root/reference):
  encode   -> modules/encode.py:63-79
  decode   -> modules/render.py:9-47
  render   -> modules/render.py:204-252
  fitness  -> modules/fitness.py:8-31
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libggs_oracle.so")

MODE_PLAIN, MODE_MASK, MODE_BOOST = 0, 1, 2

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile libggs_oracle.so with the Makefile next to this file."""
    src = os.path.join(_HERE, "ggs_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libggs_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.ggs_oracle_threads.restype = ctypes.c_int
        L.ggs_oracle_set_threads.argtypes = [ctypes.c_int]
        L.ggs_oracle_encode.argtypes = [_f32p, _f32p, ctypes.c_int64, ctypes.c_int]
        L.ggs_oracle_decode.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_float, _f32p, _i32p]
        L.ggs_oracle_render.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_float, _f32p, _f32p]
        L.ggs_oracle_render.restype = ctypes.c_int64
        L.ggs_oracle_fitness.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_float, _f32p, _f32p,
                                         ctypes.c_int, ctypes.c_float, _f32p, _f32p, _i64p]
        L.ggs_oracle_fitness.restype = ctypes.c_int
        L.ggs_oracle_score.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p,
                                       _f32p, ctypes.c_int, ctypes.c_float, _f32p]
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a: np.ndarray | None):
    if a is None:
        return ctypes.cast(None, _f32p)
    return a.ctypes.data_as(_f32p)


def threads() -> int:
    return int(lib().ggs_oracle_threads())


def set_threads(n: int) -> None:
    lib().ggs_oracle_set_threads(int(n))


def encode(axes) -> np.ndarray:
    """[..., C>=9] axes-angle -> [..., 9] Cholesky layout (encode.py:63-79)."""
    a = _f32(axes)
    cols = a.shape[-1]
    assert cols >= 9
    rows = a.size // cols
    out = np.empty(a.shape[:-1] + (9,), dtype=np.float32)
    lib().ggs_oracle_encode(_ptr(a), _ptr(out), rows, cols)
    return out


DECODE_FLOAT_KEYS = ("cx", "cy", "sxx", "sxy", "syy", "rc", "gc", "bc", "a")
DECODE_INT_KEYS = ("x0", "x1", "y0", "y1")


def decode(chol, H: int, W: int, k_sigma: float = 3.0) -> dict:
    """[..., C>=9] Cholesky rows -> dict of the reference's 13 arrays (render.py:9-47)."""
    g = _f32(chol)
    cols = g.shape[-1]
    rows = g.size // cols
    of = np.empty((9, rows), dtype=np.float32)
    oi = np.empty((4, rows), dtype=np.int32)
    lib().ggs_oracle_decode(_ptr(g), rows, cols, H, W, k_sigma, _ptr(of),
                            oi.ctypes.data_as(_i32p))
    shape = g.shape[:-1]
    out = {k: of[i].reshape(shape) for i, k in enumerate(DECODE_FLOAT_KEYS)}
    out.update({k: oi[i].reshape(shape) for i, k in enumerate(DECODE_INT_KEYS)})
    return out


def render(chol, H: int, W: int, k_sigma: float = 3.0, background=(1.0, 1.0, 1.0),
           return_pairs: bool = False):
    """[B,N,C] or [N,C] Cholesky genomes -> [B,H,W,3] float32 (render.py:204-252)."""
    g = _f32(chol)
    if g.ndim == 2:
        g = g[None]
    assert g.ndim == 3 and g.shape[2] >= 9
    B, N, C = g.shape
    bg = _f32(background)
    img = np.empty((B, H, W, 3), dtype=np.float32)
    pairs = lib().ggs_oracle_render(_ptr(g), B, N, C, H, W, k_sigma, _ptr(bg), _ptr(img))
    return (img, int(pairs)) if return_pairs else img


def _mode(weight_mask, boost_only: bool) -> int:
    if weight_mask is None:
        return MODE_PLAIN
    return MODE_BOOST if boost_only else MODE_MASK


def fitness(axes, target, H: int, W: int, k_sigma: float = 3.0, weight_mask=None,
            boost_only: bool = False, boost_beta: float = 1.0, return_images: bool = False,
            return_pairs: bool = False):
    """[B,N,C] axes-angle genomes -> [B] float32 fitness (fitness.py:8-31)."""
    g = _f32(axes)
    if g.ndim == 2:
        g = g[None]
    B, N, C = g.shape
    t = _f32(target)
    assert t.shape == (H, W, 3)
    m = None if weight_mask is None else _f32(weight_mask)
    if m is not None:
        assert m.shape == (H, W)
    fit = np.empty((B,), dtype=np.float32)
    img = np.empty((B, H, W, 3), dtype=np.float32) if return_images else None
    pairs = ctypes.c_int64(0)
    rc = lib().ggs_oracle_fitness(_ptr(g), B, N, C, H, W, k_sigma, _ptr(t), _ptr(m),
                                  _mode(m, boost_only), boost_beta, _ptr(fit), _ptr(img),
                                  ctypes.byref(pairs))
    if rc != 0:
        raise RuntimeError(f"ggs_oracle_fitness failed: {rc}")
    out = [fit]
    if return_images:
        out.append(img)
    if return_pairs:
        out.append(int(pairs.value))
    return out[0] if len(out) == 1 else tuple(out)


def score(images, target, weight_mask=None, boost_only: bool = False, boost_beta: float = 1.0):
    """Score rendered images [B,H,W,3] (fitness.py:16-31)."""
    im = _f32(images)
    B, H, W, _ = im.shape
    t = _f32(target)
    m = None if weight_mask is None else _f32(weight_mask)
    fit = np.empty((B,), dtype=np.float32)
    lib().ggs_oracle_score(_ptr(im), B, H, W, _ptr(t), _ptr(m), _mode(m, boost_only), boost_beta,
                           _ptr(fit))
    return fit
