"""Channel sharding across GPUs (one process per GPU, ``torch.distributed``).

The reference has no multi-device code at all (no NCCL / torch.distributed
call sites).  Channels and filterbank lanes are independent, so the data path needs NO
collective: rank r of P filters channels ``[r*C/P, (r+1)*C/P)`` with its own state
(SURVEY.md 8e).  The only collective is an optional all-gather of the output block when a
caller wants the whole ``[C, T]`` on every rank (``north_star``: "an NCCL all-gather only
to reassemble the multichannel output"); with the channel-major layout every rank's block
is contiguous, so the gather needs no repacking.  Works with the ``nccl`` backend on GPUs
and ``gloo`` on CPU (tests/test_dist.py runs world_size 2 on gloo).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import Tensor, nn


def _world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_bounds(num_channels: int, world_size: int, rank: int) -> tuple[int, int]:
    """Half-open channel range owned by ``rank``: sizes differ by at most one, lower ranks
    take the larger blocks, the union is exactly ``[0, num_channels)``."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, extra = divmod(num_channels, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_channels(x: Tensor, group=None, dim: int = -2) -> Tensor:
    """This rank's channel block of ``x`` (a view, no copy)."""
    world, rank = _world(group)
    lo, hi = shard_bounds(x.shape[dim], world, rank)
    return x.narrow(dim, lo, hi - lo)


def all_gather_channels(y_local: Tensor, num_channels: int, group=None, out: Tensor | None = None) -> Tensor:
    """Reassemble ``[C, T]`` from the per-rank blocks ``[C_r, T]`` (rank order == channel order)."""
    world, rank = _world(group)
    T = y_local.shape[-1]
    if out is None:
        out = torch.empty((num_channels, T), dtype=y_local.dtype, device=y_local.device)
    if world == 1:
        out.copy_(y_local)
        return out
    y_local = y_local.contiguous()
    if num_channels % world == 0:
        dist.all_gather_into_tensor(out, y_local, group=group)
        return out
    # ragged: one broadcast-sized slot per rank, then trim
    per = -(-num_channels // world)
    slot = torch.zeros((per, T), dtype=y_local.dtype, device=y_local.device)
    slot[: y_local.shape[0]] = y_local
    slots = torch.empty((world * per, T), dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(slots, slot, group=group)
    for r in range(world):
        lo, hi = shard_bounds(num_channels, world, r)
        out[lo:hi] = slots[r * per : r * per + (hi - lo)]
    return out


class ChannelSharded(nn.Module):
    """Run ``module`` on this rank's channel block of a ``[C, T]`` input.

    ``gather=False`` (default) returns the local block -- no collective on the data path.
    ``gather=True`` all-gathers the blocks so every rank returns the full ``[C, T]``.
    The wrapped filter's state is rank-local (``[K, C/P, 2]``), coefficients are replicated.
    """

    def __init__(self, module: nn.Module, group=None, gather: bool = False, input_is_sharded: bool = False) -> None:
        super().__init__()
        self.module = module
        self.group = group
        self.gather = gather
        self.input_is_sharded = input_is_sharded

    def forward(self, x: Tensor, num_channels: int | None = None) -> Tensor:
        if self.input_is_sharded:
            local = x
            if self.gather and num_channels is None:
                raise ValueError("num_channels is required to gather a pre-sharded input")
            total = num_channels
        else:
            total = x.shape[-2]
            local = shard_channels(x, self.group)
        y = self.module(local)
        if not self.gather:
            return y
        return all_gather_channels(y, int(total), self.group)
