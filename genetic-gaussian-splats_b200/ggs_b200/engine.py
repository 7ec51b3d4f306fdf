"""GA generations enqueued back to back (ggs_ga_* in include/ggs_b200.h): breeding, evaluation,
elitism, ranking and curve statistics all stay on the device; the host reads results when it
wants them.  Counterpart of the generation loop of the reference's modules/algorithm.py:87-160.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from .evaluator import SIGMA_ORDER, _as_f32, _cuda_device, _stream_ptr
from .native import MODE_BOOST, MODE_MASK, MODE_PLAIN, check, lib

MAX_POPULATION = 16384   # the ranking kernel sorts the population in one CTA's shared memory


class _DeviceView:
    """Zero-copy torch view of engine-owned device memory (__cuda_array_interface__)."""

    def __init__(self, ptr: int, shape: Tuple[int, ...]):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False),
                                         "version": 2, "strides": None}


class GaEngine:
    """One GA run on one device.

        eng = GaEngine(target, mask, H, W, P, N, n_elite, max_generations)
        eng.start(population, seed)                 # generation 0
        eng.run(sigma_rows, tour_k, cxpb, mutpb, log_lo, log_hi)   # enqueue len(sigma_rows) generations
        st = eng.state()                            # sync; curves, best individual
    """

    def __init__(self, target: torch.Tensor, weight_mask: Optional[torch.Tensor], H: int, W: int,
                 P: int, N: int, n_elite: int, max_generations: int, *, k_sigma: float = 3.0,
                 boost_only: bool = False, boost_beta: float = 1.0, device=None):
        self.device = _cuda_device(device if device is not None else target.device)
        self.P, self.N, self.H, self.W = int(P), int(N), int(H), int(W)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ggs_ga_create(self.device.index or 0, self.P, self.N, self.H, self.W,
                                      int(n_elite), int(max_generations), ctypes.byref(self._h)),
                  "ggs_ga_create")
            t = _as_f32(target, self.device)
            assert t.shape == (self.H, self.W, 3), "target must be [H, W, 3]"
            m = None if weight_mask is None else _as_f32(weight_mask, self.device)
            assert m is None or m.shape == (self.H, self.W), "weight_mask must be [H, W]"
            mode = MODE_PLAIN if m is None else (MODE_BOOST if boost_only else MODE_MASK)
            check(lib().ggs_ga_set_target(self._h, t.data_ptr(), None if m is None else m.data_ptr(),
                                          mode, float(boost_beta), float(k_sigma),
                                          _stream_ptr(self.device)), "ggs_ga_set_target")
            torch.cuda.current_stream(self.device).synchronize()   # t / m may be temporaries

    def set_peers(self, peers) -> None:
        """Shard the evaluation over the ranks of a ggs_b200.peers.PeerGroup (before start()):
        every rank runs the same calls on the same population and seed, evaluates its slice of
        the children, and the fitness values travel GPU to GPU inside the raster kernel."""
        self._peers = peers     # keep it alive as long as the engine
        check(lib().ggs_ga_set_peers(self._h, None if peers is None else peers._h), "ggs_ga_set_peers")

    def start(self, population: torch.Tensor, seed: int) -> None:
        pop = _as_f32(population, self.device)
        assert pop.shape[:2] == (self.P, self.N) and pop.shape[2] >= 9
        with torch.cuda.device(self.device):
            check(lib().ggs_ga_start(self._h, pop.data_ptr(), int(pop.shape[2]),
                                     int(seed) & (2**64 - 1), _stream_ptr(self.device)), "ggs_ga_start")
            torch.cuda.current_stream(self.device).synchronize()   # `pop` may be a temporary

    def run(self, sigma_rows: Sequence[dict], tour_k: int, cxpb: float, mutpb: float,
            log_scale_lo: float, log_scale_hi: float) -> None:
        """Enqueue one generation per row of `sigma_rows` (dicts keyed like MUT_SIGMA_MAX)."""
        count = len(sigma_rows)
        if count == 0:
            return
        flat = (ctypes.c_float * (6 * count))(*[float(r[k]) for r in sigma_rows for k in SIGMA_ORDER])
        with torch.cuda.device(self.device):
            check(lib().ggs_ga_run(self._h, count, flat, int(tour_k), float(cxpb), float(mutpb),
                                   float(log_scale_lo), float(log_scale_hi),
                                   _stream_ptr(self.device)), "ggs_ga_run")

    def state(self, curves_from: int = 0, want_best: bool = True) -> dict:
        """Synchronise and read: generation, best fitness, stale count, curve points
        [curves_from ..] as an [k, 3] array (best so far, mean, median), best individual."""
        gen, stale, best = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        ind = np.empty((self.N, 9), dtype=np.float32) if want_best else None
        # the curve buffer must hold every point up to the generation reached by queued work
        done = ctypes.c_int()
        with torch.cuda.device(self.device):
            check(lib().ggs_ga_state(self._h, _stream_ptr(self.device), ctypes.byref(done), None, None,
                                     None, 0, None), "ggs_ga_state")
            curves = np.empty((max(done.value + 1 - curves_from, 0), 3), dtype=np.float64)
            check(lib().ggs_ga_state(self._h, _stream_ptr(self.device), ctypes.byref(gen),
                                     ctypes.byref(best), ctypes.byref(stale),
                                     curves.ctypes.data if curves.size else None, int(curves_from),
                                     None if ind is None else ind.ctypes.data), "ggs_ga_state")
        return {"generation": gen.value, "best_fitness": best.value, "no_improve": stale.value,
                "curves": curves,
                "best_individual": None if ind is None else torch.from_numpy(ind)}

    def population(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Copies of the current population [P,N,9] and its fitness [P] (device tensors)."""
        p, f = ctypes.c_void_p(), ctypes.c_void_p()
        check(lib().ggs_ga_population(self._h, ctypes.byref(p), ctypes.byref(f)), "ggs_ga_population")
        with torch.cuda.device(self.device):
            pop = torch.as_tensor(_DeviceView(p.value, (self.P, self.N, 9)), device=self.device).clone()
            fit = torch.as_tensor(_DeviceView(f.value, (self.P,)), device=self.device).clone()
        return pop, fit

    def close(self) -> None:
        if self._h:
            lib().ggs_ga_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


MAX_TRIES = 64   # ggs_sa_*: the Metropolis kernel takes one U[0,1) draw per try as a launch argument


class SaEngine:
    """One simulated-annealing chain on one device (ggs_sa_* in include/ggs_b200.h): the
    iteration loop of the reference's modules/annealing.py:112-150 with batched neighbours.

        eng = SaEngine(target, mask, H, W, N, tries, max_iterations)
        eng.start(state, seed)                                    # iteration 0
        eng.run(sigma_rows, temperatures, uniforms, mutpb, lo, hi) # enqueue len(sigma_rows) iterations
        st = eng.state()                                          # sync; curves, best / current state
    """

    def __init__(self, target: torch.Tensor, weight_mask: Optional[torch.Tensor], H: int, W: int,
                 N: int, tries: int, max_iterations: int, *, k_sigma: float = 3.0,
                 boost_only: bool = False, boost_beta: float = 1.0, device=None,
                 batch_neighbors: bool = False):
        """batch_neighbors=False: the reference's chain, every try starts from the state the
        previous one left (annealing.py:121-146); True: all tries of an iteration are proposed
        from the same state and scored by one evaluation (BASELINE config 2)."""
        self.device = _cuda_device(device if device is not None else target.device)
        self.N, self.H, self.W, self.tries = int(N), int(H), int(W), int(tries)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ggs_sa_create(self.device.index or 0, self.N, self.H, self.W, self.tries,
                                      int(max_iterations), ctypes.byref(self._h)), "ggs_sa_create")
            check(lib().ggs_sa_set_mode(self._h, 1 if batch_neighbors else 0), "ggs_sa_set_mode")
            t = _as_f32(target, self.device)
            assert t.shape == (self.H, self.W, 3), "target must be [H, W, 3]"
            m = None if weight_mask is None else _as_f32(weight_mask, self.device)
            assert m is None or m.shape == (self.H, self.W), "weight_mask must be [H, W]"
            mode = MODE_PLAIN if m is None else (MODE_BOOST if boost_only else MODE_MASK)
            check(lib().ggs_sa_set_target(self._h, t.data_ptr(), None if m is None else m.data_ptr(),
                                          mode, float(boost_beta), float(k_sigma),
                                          _stream_ptr(self.device)), "ggs_sa_set_target")
            torch.cuda.current_stream(self.device).synchronize()   # t / m may be temporaries

    def start(self, state: torch.Tensor, seed: int) -> None:
        s = _as_f32(state, self.device)
        assert s.dim() == 2 and s.shape[0] == self.N and s.shape[1] >= 9
        with torch.cuda.device(self.device):
            check(lib().ggs_sa_start(self._h, s.data_ptr(), int(s.shape[1]), int(seed) & (2**64 - 1),
                                     _stream_ptr(self.device)), "ggs_sa_start")
            torch.cuda.current_stream(self.device).synchronize()   # `s` may be a temporary

    def run(self, sigma_rows: Sequence[dict], temperatures: Sequence[float],
            uniforms: Sequence[Sequence[float]], mutpb: float, log_scale_lo: float,
            log_scale_hi: float) -> None:
        """Enqueue one iteration per row: its annealed sigmas (dict keyed like MUT_SIGMA_MAX), its
        temperature and `tries` U[0,1) draws for the uphill tests."""
        count = len(sigma_rows)
        if count == 0:
            return
        assert len(temperatures) == count and len(uniforms) == count
        sig = (ctypes.c_float * (6 * count))(*[float(r[k]) for r in sigma_rows for k in SIGMA_ORDER])
        temp = (ctypes.c_double * count)(*[float(t) for t in temperatures])
        uni = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64).reshape(count, self.tries))
        with torch.cuda.device(self.device):
            check(lib().ggs_sa_run(self._h, count, sig, temp,
                                   uni.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), float(mutpb),
                                   float(log_scale_lo), float(log_scale_hi),
                                   _stream_ptr(self.device)), "ggs_sa_run")

    def state(self, curves_from: int = 0, want_states: bool = True) -> dict:
        """Synchronise and read: iteration, best / current energy, curve points [curves_from ..]
        as a [k, 2] array (best, current), the best and the current state."""
        it, e_best, e_cur = ctypes.c_int(), ctypes.c_double(), ctypes.c_double()
        best = np.empty((self.N, 9), dtype=np.float32) if want_states else None
        cur = np.empty((self.N, 9), dtype=np.float32) if want_states else None
        done = ctypes.c_int()
        with torch.cuda.device(self.device):
            check(lib().ggs_sa_state(self._h, _stream_ptr(self.device), ctypes.byref(done), None, None,
                                     None, 0, None, None), "ggs_sa_state")
            curves = np.empty((max(done.value + 1 - curves_from, 0), 2), dtype=np.float64)
            check(lib().ggs_sa_state(self._h, _stream_ptr(self.device), ctypes.byref(it),
                                     ctypes.byref(e_best), ctypes.byref(e_cur),
                                     curves.ctypes.data if curves.size else None, int(curves_from),
                                     None if best is None else best.ctypes.data,
                                     None if cur is None else cur.ctypes.data), "ggs_sa_state")
        return {"iteration": it.value, "best_energy": e_best.value, "current_energy": e_cur.value,
                "curves": curves,
                "best_state": None if best is None else torch.from_numpy(best),
                "current_state": None if cur is None else torch.from_numpy(cur)}

    def close(self) -> None:
        if self._h:
            lib().ggs_sa_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
