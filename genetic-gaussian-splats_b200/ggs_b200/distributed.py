"""Population sharding over the GPUs of one box: one process per GPU, `torch.distributed`.

The render + fitness path shards naturally: candidates are independent,
so rank r evaluates the contiguous slice [lo_r, hi_r) of the population with the same kernel and
one all-gather moves the fp32 fitness vector (4 B per candidate).  No genome traffic is needed
when every rank breeds the same next generation from the same seed ("replicated deterministic
breeding"); `gather_rows` covers the other case (each rank only owns its shard) by moving just
the selected individuals (the elites, algorithm.py:128-131).

Because the per-candidate reduction order inside the kernel is fixed, the gathered vector is
bit-identical to a single-GPU evaluation of the whole population.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env() -> Tuple[int, int]:
    """(rank, world) of this process.  Under `torchrun` (WORLD_SIZE > 1 in the environment) the
    process binds to cuda:LOCAL_RANK and joins the NCCL group on first use, so the reference's
    run scripts shard a GA run over the GPUs of a box without any change:
        torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 run_ggs.py
    GGS_B200_NO_SHARD=1 keeps a process on its own (every rank then runs the whole job)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or os.environ.get("GGS_B200_NO_SHARD", "0") == "1":
        return 0, 1
    if not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank(), dist.get_world_size()


def replicate(t: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """Rank `src`'s copy of `t` on every rank (in place): makes the state that replicated
    deterministic breeding starts from identical even when the ranks' RNGs were not seeded alike."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return t


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of `total` items; the first `total % world` ranks get one extra."""
    assert world >= 1 and 0 <= rank < world and total >= 0
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0]
            for r in range(world)]


def owner_of(index: int, total: int, world: int) -> int:
    for r in range(world):
        lo, hi = shard_bounds(total, world, r)
        if lo <= index < hi:
            return r
    raise IndexError(index)


class ShardedEvaluator:
    """fitness_many over a population sharded across the ranks of a process group.

    evaluate(genomes[b,N,C]) -> Tensor[b] is the per-rank evaluation; by default the fused
    CUDA path (`ggs_b200.fitness`) with the target / mask resident on this rank's device.  Tests
    inject a CPU function to exercise the sharding logic under gloo.
    """

    def __init__(self, target: torch.Tensor, H: int, W: int, k_sigma: float = 3.0,
                 weight_mask: Optional[torch.Tensor] = None, boost_only: bool = False,
                 device=None, group=None,
                 evaluate: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 peers=None):
        self.group = group
        self._takes_total = False
        # peers: a ggs_b200.peers.PeerGroup -> the raster kernel delivers the fitness values to
        # every rank itself (no collective launch); None -> one NCCL / gloo all-gather
        self.peers = peers
        self._p2p = None
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.H, self.W, self.k_sigma = int(H), int(W), float(k_sigma)
        if evaluate is None:
            from . import evaluator
            dev = torch.device(device if device is not None else "cuda")
            tgt = target.to(dev, torch.float32).contiguous()
            msk = None if weight_mask is None else weight_mask.to(dev, torch.float32).contiguous()

            def evaluate(g: torch.Tensor, total: Optional[int] = None) -> torch.Tensor:
                # the kernel configuration of the WHOLE population, so that the gathered vector
                # has the bits of a single-GPU evaluation
                split = evaluator.choose_split(total or g.shape[0], g.shape[1], self.H, self.W)
                return evaluator.fitness(g, tgt, self.H, self.W, self.k_sigma, weight_mask=msk,
                                         boost_only=boost_only, device=dev, split=split)
            self._takes_total = True
            if peers is not None:
                def p2p(shard: torch.Tensor, lo: int, total: int) -> torch.Tensor:
                    return peers.fitness_allgather(shard, tgt, self.H, self.W, offset=lo, total=total,
                                                   k_sigma=self.k_sigma, weight_mask=msk,
                                                   boost_only=boost_only)
                self._p2p = p2p
        self.evaluate = evaluate

    # -- fitness -------------------------------------------------------------------------
    def local_slice(self, total: int) -> slice:
        lo, hi = shard_bounds(total, self.world, self.rank)
        return slice(lo, hi)

    @torch.no_grad()
    def fitness(self, population: torch.Tensor, *, replicated: bool = True,
                total: Optional[int] = None) -> torch.Tensor:
        """Fitness of the WHOLE population on every rank.

        replicated=True : `population` is the full [P,N,C] tensor, identical on all ranks;
                          this rank evaluates only its slice.
        replicated=False: `population` is this rank's shard; `total` = P.
        """
        if replicated:
            P = population.shape[0]
            shard = population[self.local_slice(P)]
        else:
            assert total is not None, "total population size required for sharded input"
            P, shard = int(total), population
            lo, hi = shard_bounds(P, self.world, self.rank)
            assert shard.shape[0] == hi - lo, "shard does not match shard_bounds()"
        if self._p2p is not None and self.world > 1:
            lo, _ = shard_bounds(P, self.world, self.rank)
            return self._p2p(shard.contiguous(), lo, P).clone()   # the view lives for two gathers
        if shard.shape[0] > 0:
            local = (self.evaluate(shard.contiguous(), P) if self._takes_total
                     else self.evaluate(shard.contiguous()))
        else:
            local = torch.empty((0,), dtype=torch.float32, device=population.device)
        if self.world == 1:
            return local
        return self._all_gather(local, P)

    def _all_gather(self, local: torch.Tensor, P: int) -> torch.Tensor:
        sizes = shard_sizes(P, self.world)
        if len(set(sizes)) == 1:
            out = torch.empty((P,), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
            return out
        # ragged: pad every shard to the largest, gather, strip the padding
        width = max(sizes)
        padded = torch.zeros((width,), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        out = torch.empty((self.world * width,), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=self.group)
        return torch.cat([out[r * width: r * width + sizes[r]] for r in range(self.world)])

    # -- elites --------------------------------------------------------------------------
    @torch.no_grad()
    def gather_rows(self, shard: torch.Tensor, indices: Sequence[int], total: int) -> torch.Tensor:
        """Individuals `indices` (global numbering) on every rank, from per-rank shards.
        Each row has exactly one owner, so a SUM all-reduce of zero-padded rows is exact."""
        lo, hi = shard_bounds(total, self.world, self.rank)
        out = torch.zeros((len(indices),) + tuple(shard.shape[1:]), dtype=shard.dtype,
                          device=shard.device)
        for k, i in enumerate(indices):
            if lo <= i < hi:
                out[k] = shard[i - lo]
        if self.world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out

    def elites(self, fitness_full: torch.Tensor, elite_k: int) -> List[int]:
        """Indices of the elite_k best (lowest) fitness values, ties by index: the same list
        on every rank (algorithm.py:128-130 sorts by fitness with a stable sort)."""
        order = torch.argsort(fitness_full.detach().cpu(), stable=True)
        return order[: max(1, int(elite_k))].tolist()
