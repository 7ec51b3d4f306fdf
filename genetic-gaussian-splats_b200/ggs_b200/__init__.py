"""ggs_b200: B200-native render + fitness hot path of genetic-gaussian-splats.

Python here is host-side plumbing over libggs_b200.so (C ABI in include/ggs_b200.h).
"""
from .native import (LAYOUT_AXES_ANGLE, LAYOUT_CHOLESKY, MODE_BOOST, MODE_MASK, MODE_PLAIN,
                     GgsError, lib)
from .evaluator import (HostEvaluator, breed, choose_split, count_evaluated_pairs, set_option, tile_order, decode, encode, fitness, importance_mask, mode_of,
                        probe_peaks, render,
                        timing_enable, timing_read)

__all__ = ["LAYOUT_AXES_ANGLE", "LAYOUT_CHOLESKY", "MODE_PLAIN", "MODE_MASK", "MODE_BOOST",
           "GgsError", "lib", "HostEvaluator", "breed", "choose_split", "set_option", "tile_order", "count_evaluated_pairs", "decode", "encode", "fitness", "importance_mask", "mode_of",
           "probe_peaks", "render", "timing_enable", "timing_read"]
