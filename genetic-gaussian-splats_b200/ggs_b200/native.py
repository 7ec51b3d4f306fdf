"""ctypes binding of libggs_b200.so (C ABI: include/ggs_b200.h).

There is no fallback: if the library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes
import os

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("GGS_B200_LIB") or os.path.join(_PKG_ROOT, "lib", "libggs_b200.so")

OK = 0
LAYOUT_AXES_ANGLE, LAYOUT_CHOLESKY = 0, 1
MODE_PLAIN, MODE_MASK, MODE_BOOST = 0, 1, 2
MAX_SIDE = 32768
ABI_VERSION = 1

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f = ctypes.c_float
_d = ctypes.c_double
_sz = ctypes.c_size_t

# name -> (restype, argtypes); one entry per symbol declared in include/ggs_b200.h
SIGNATURES = {
    "ggs_abi_version": (_i, []),
    "ggs_last_error": (ctypes.c_char_p, []),
    "ggs_device_count": (_i, []),
    "ggs_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ggs_encode": (_i, [_vp, _i64, _i, _vp, _vp]),
    "ggs_decode": (_i, [_vp, _i, _i64, _i, _i, _i, _f, _vp, _vp, _vp]),
    "ggs_render": (_i, [_vp, _i, _i, _i, _i, _i, _i, _f, ctypes.POINTER(_f), _vp, _vp, _sz, _vp]),
    "ggs_render_u8": (_i, [_vp, _i, _i, _i, _i, _i, _i, _f, ctypes.POINTER(_f), _vp, _vp, _sz, _vp]),
    "ggs_fitness": (_i, [_vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _i, _f, _vp, _vp, _vp, _sz,
                         _vp]),
    "ggs_fitness_ex": (_i, [_vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _i, _f, _vp, _vp, _vp, _sz,
                            _i, _vp]),
    "ggs_choose_split": (_i, [_i, _i, _i, _i]),
    "ggs_set_option": (_i, [ctypes.c_char_p, _i]),
    "ggs_tile_order": (_i, [_i, _i, ctypes.POINTER(_i)]),
    "ggs_ctx_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "ggs_ctx_destroy": (None, [_vp]),
    "ggs_ctx_set_target": (_i, [_vp, _vp, _vp, _i, _i]),
    "ggs_ctx_fitness_host": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _f, _vp]),
    "ggs_probe_peaks": (_i, [ctypes.POINTER(_f)]),
    "ggs_ga_breed": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _f, _f, ctypes.POINTER(_f), _f, _f,
                          ctypes.c_uint64, ctypes.c_uint32, _vp]),
    "ggs_ga_create": (_i, [_i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_vp)]),
    "ggs_ga_destroy": (None, [_vp]),
    "ggs_ga_set_target": (_i, [_vp, _vp, _vp, _i, _f, _f, _vp]),
    "ggs_ga_start": (_i, [_vp, _vp, _i, ctypes.c_uint64, _vp]),
    "ggs_ga_run": (_i, [_vp, _i, ctypes.POINTER(_f), _i, _f, _f, _f, _f, _vp]),
    "ggs_ga_state": (_i, [_vp, _vp, ctypes.POINTER(_i), ctypes.POINTER(_d), ctypes.POINTER(_i), _vp,
                          _i, _vp]),
    "ggs_ga_population": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "ggs_sa_create": (_i, [_i, _i, _i, _i, _i, _i, ctypes.POINTER(_vp)]),
    "ggs_sa_destroy": (None, [_vp]),
    "ggs_sa_set_mode": (_i, [_vp, _i]),
    "ggs_sa_set_target": (_i, [_vp, _vp, _vp, _i, _f, _f, _vp]),
    "ggs_sa_start": (_i, [_vp, _vp, _i, ctypes.c_uint64, _vp]),
    "ggs_sa_run": (_i, [_vp, _i, ctypes.POINTER(_f), ctypes.POINTER(_d), ctypes.POINTER(_d), _f, _f,
                        _f, _vp]),
    "ggs_sa_state": (_i, [_vp, _vp, ctypes.POINTER(_i), ctypes.POINTER(_d), ctypes.POINTER(_d), _vp,
                          _i, _vp, _vp]),
    "ggs_peers_create": (_i, [_i, _i, _i, _i, ctypes.POINTER(_vp)]),
    "ggs_peers_destroy": (None, [_vp]),
    "ggs_peers_export": (_i, [_vp, _vp]),
    "ggs_peers_connect": (_i, [_vp, _vp]),
    "ggs_peers_connect_local": (_i, [_vp, ctypes.POINTER(_vp)]),
    "ggs_peers_status": (_i, [_vp, _vp]),
    "ggs_fitness_allgather": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _i, _f, _i, _i, _vp,
                                   _sz, ctypes.POINTER(_vp), _vp]),
    "ggs_ga_set_peers": (_i, [_vp, _vp]),
    "ggs_mask_workspace_bytes": (_sz, [_i, _i]),
    "ggs_importance_mask": (_i, [_vp, _i, _i, _i, _i, _i, ctypes.POINTER(_i), _i, _d, _d, _d, _d, _i,
                                 _d, _vp, _vp, _sz, _vp]),
    "ggs_stats_target": (_i, [_vp]),
    "ggs_timing_enable": (_i, [_i]),
    "ggs_timing_read": (_i, [ctypes.POINTER(_f), ctypes.POINTER(_f), ctypes.POINTER(_i)]),
}


class GgsError(RuntimeError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    """Load libggs_b200.so once; raise loudly when it is missing (no CPU fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GgsError(
                f"{LIB_PATH} not found: build it with "
                f"`python genetic-gaussian-splats_b200/build.py` (or __graft_entry__.build()). "
                "There is no CPU fallback for the render/fitness path.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        got = L.ggs_abi_version()
        if got != ABI_VERSION:
            raise GgsError(f"libggs_b200.so ABI {got} != binding ABI {ABI_VERSION}")
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = lib().ggs_last_error()
        raise GgsError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
