"""The fitness all-gather between the GPUs of one box without a collective launch
(ggs_peers_* in include/ggs_b200.h): the raster kernel stores every fitness value straight into
each rank's receive buffer over NVLink and raises a flag; consumers wait for the flags.

PyTorch is plumbing: `torch.distributed` carries the 64-byte CUDA IPC handles once, at set-up.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from .engine import _DeviceView
from .evaluator import _as_f32, _cuda_device, _stream_ptr, _workspace, mode_of
from .native import LAYOUT_AXES_ANGLE, check, lib

IPC_HANDLE_BYTES = 64
MAX_WORLD = 8


class PeerGroup:
    """One rank's end of the exchange.

        peers = PeerGroup.from_process_group(capacity=P)          # one process per GPU (torchrun)
        full = peers.fitness_allgather(shard, target, H, W, offset=lo, total=P, weight_mask=m)

    `full` is a view of this rank's gathered vector [total]: valid in stream order, until the call
    after the next one (two epochs are buffered)."""

    def __init__(self, rank: int, world: int, capacity: int, device=None):
        self.device = _cuda_device(device)
        self.rank, self.world, self.capacity = int(rank), int(world), int(capacity)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ggs_peers_create(self.device.index or 0, self.rank, self.world, self.capacity,
                                         ctypes.byref(self._h)), "ggs_peers_create")

    # ---- set-up ------------------------------------------------------------------------
    def export_handle(self) -> bytes:
        buf = (ctypes.c_char * IPC_HANDLE_BYTES)()
        check(lib().ggs_peers_export(self._h, buf), "ggs_peers_export")
        return bytes(buf)

    def connect(self, handles: Sequence[bytes]) -> None:
        assert len(handles) == self.world and all(len(h) == IPC_HANDLE_BYTES for h in handles)
        blob = b"".join(handles)
        with torch.cuda.device(self.device):
            check(lib().ggs_peers_connect(self._h, blob), "ggs_peers_connect")

    @classmethod
    def from_process_group(cls, capacity: int, device=None, group=None) -> "PeerGroup":
        """Collective over `group` (NCCL): every rank creates its end, the IPC handles are
        all-gathered, every rank maps the others."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        assert world <= MAX_WORLD, f"at most {MAX_WORLD} ranks (one box)"
        self = cls(rank, world, capacity, device)
        mine = torch.tensor(list(self.export_handle()), dtype=torch.uint8, device=self.device)
        everyone = torch.empty((world * IPC_HANDLE_BYTES,), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(everyone, mine, group=group)
        blob = bytes(everyone.cpu().tolist())
        self.connect([blob[r * IPC_HANDLE_BYTES:(r + 1) * IPC_HANDLE_BYTES] for r in range(world)])
        dist.barrier(group=group)
        return self

    @staticmethod
    def connect_local(groups: List["PeerGroup"]) -> None:
        """Ranks living in one process: connect them through the objects themselves."""
        arr = (ctypes.c_void_p * len(groups))(*[g._h for g in groups])
        for g in groups:
            with torch.cuda.device(g.device):
                check(lib().ggs_peers_connect_local(g._h, arr), "ggs_peers_connect_local")

    # ---- the gather --------------------------------------------------------------------
    @torch.no_grad()
    def fitness_allgather(self, shard: torch.Tensor, target: torch.Tensor, H: int, W: int, *,
                          offset: int, total: int, k_sigma: float = 3.0,
                          weight_mask: Optional[torch.Tensor] = None, boost_only: bool = False,
                          boost_beta: float = 1.0, layout: int = LAYOUT_AXES_ANGLE) -> torch.Tensor:
        """Evaluate this rank's candidates [offset, offset + B) of a population of `total`; returns
        this rank's view of the whole fitness vector [total] once every rank's values are in."""
        dev = self.device
        g = _as_f32(shard, dev)
        assert g.ndim == 3 and g.shape[2] >= 9
        B, N, C = g.shape
        t = _as_f32(target, dev)
        m = None if weight_mask is None else _as_f32(weight_mask, dev)
        out = ctypes.c_void_p()
        with torch.cuda.device(dev):
            ws = _workspace(dev, lib().ggs_workspace_bytes(max(B, 1), N, int(H), int(W)))
            check(lib().ggs_fitness_allgather(
                self._h, g.data_ptr() if B else None, layout, B, N, C, int(H), int(W), float(k_sigma),
                t.data_ptr(), None if m is None else m.data_ptr(), mode_of(m, boost_only),
                float(boost_beta), int(offset), int(total), ws.data_ptr(), ws.numel(),
                ctypes.byref(out), _stream_ptr(dev)), "ggs_fitness_allgather")
            return torch.as_tensor(_DeviceView(out.value, (int(total),)), device=dev)

    def check(self) -> None:
        """Synchronise and raise if a wait for a peer timed out."""
        with torch.cuda.device(self.device):
            check(lib().ggs_peers_status(self._h, _stream_ptr(self.device)), "ggs_peers_status")

    def close(self) -> None:
        if self._h:
            lib().ggs_peers_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
