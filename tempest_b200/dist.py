"""Particle sharding over the GPUs of one node.

One process per GPU (torchrun).  Walker slots are block-partitioned: rank r owns slots
[r*N/G, (r+1)*N/G) and appends its N/G mutated particles to its local history shard every
generation, so the persistent ensemble is sharded identically; per-generation scalars and the
mode statistics are replicated.  The data path never moves the ensemble: the collectives are
(i) an all-gather of one (max, S1, S2) triple per ESS probe, (ii) all-reduces of histograms /
moment sums (<= a few thousand numbers), (iii) an ownership-masked all-reduce of the N resampled
rows (each slot's row is written by exactly one rank, every other rank contributes zeros, so the
sum is exact and order-independent) and (iv) one (K+3)-number all-reduce per Metropolis step.
Pure host logic lives here so that it can be exercised with the gloo backend on CPU.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


class Comm:
    """Thin wrapper: degenerates to no-ops for a single process."""

    def __init__(self, group=None):
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.world = dist.get_world_size(group) if self.on else 1
        self.rank = dist.get_rank(group) if self.on else 0

    def allreduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.on:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        """[world, *t.shape] in rank order."""
        if not self.on:
            return t.unsqueeze(0)
        flat = t.contiguous().reshape(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.reshape((self.world,) + tuple(t.shape))

    def allgather_rows(self, t: torch.Tensor) -> torch.Tensor:
        """Concatenate variable-length first dimensions in rank order (padded all-gather)."""
        if not self.on:
            return t
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        counts = self.allgather(n).flatten().tolist()
        cap = max(counts)
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        allp = self.allgather(pad)
        return torch.cat([allp[r, : counts[r]] for r in range(self.world)], dim=0)


def merge_ess_triples(triples: Sequence[Sequence[float]]) -> Tuple[float, float, float]:
    """Fold per-shard (m, S1, S2) of w = exp(a - m) in rank order (same rule as tb::Ess3::merge)."""
    m, s1, s2 = -math.inf, 0.0, 0.0
    for m2, a2, b2 in triples:
        if m2 == -math.inf:
            continue
        if m == -math.inf:
            m, s1, s2 = m2, a2, b2
        elif m2 <= m:
            r = math.exp(m2 - m)
            s1 += a2 * r
            s2 += b2 * (r * r)
        else:
            r = math.exp(m - m2)
            s1 = s1 * r + a2
            s2 = s2 * (r * r) + b2
            m = m2
    return m, s1, s2


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Slot range of `rank` (block partition; n must divide evenly so shards stay congruent)."""
    if n % world:
        raise ValueError(f"n_particles ({n}) must be divisible by the number of GPUs ({world})")
    per = n // world
    return rank * per, (rank + 1) * per


def exchange_owned_rows(comm: Comm, rows: torch.Tensor, owned_slots: torch.Tensor, n_global: int,
                        lo: int, hi: int) -> torch.Tensor:
    """Ownership-masked all-reduce: `rows[i]` is the payload for global slot `owned_slots[i]`
    (each slot is owned by exactly one rank).  Returns the rows of slots [lo, hi)."""
    full = torch.zeros((n_global, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    if owned_slots.numel():
        full[owned_slots] = rows
    comm.allreduce_sum_(full)
    return full[lo:hi].contiguous()
