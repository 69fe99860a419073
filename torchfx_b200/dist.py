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


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated
    afterwards (first touch) and the threads that drive the copies are local to the GPU's PCIe root.

    One process per GPU under ``torchrun`` starts unbound: with 8 ranks streaming host buffers at once, half of
    the traffic otherwise crosses the socket interconnect (round 1: e2e 11.0 -> 16.2 Gsamples/s from 1 to 8 GPUs).
    Returns what was done (``{"numa_node": n, "cpus": k}``) or why nothing was (``{"skipped": reason}``); never raises."""
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        with open(f"{base}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"skipped": "the platform reports no NUMA node for the GPU", "pci": bus}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus: set[int] = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return {"skipped": "no allowed CPU on the GPU's NUMA node", "numa_node": node, "pci": bus}
        os.sched_setaffinity(0, use)
        return {"numa_node": node, "cpus": len(use), "pci": bus}
    except Exception as e:  # pragma: no cover - depends on the host
        return {"skipped": f"{type(e).__name__}: {e}"}


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


def _carries_state(module: nn.Module) -> bool:
    """True for the SOS filters whose DF1 state makes chunk-by-chunk calls equal one call."""
    from .filter.biquad import Biquad
    from .filter.fused import FusedSOSCascade
    from .filter.iir import IIR

    return isinstance(module, (IIR, Biquad, FusedSOSCascade))


_comm_streams: dict[int, "torch.cuda.Stream"] = {}


def filter_and_gather(module: nn.Module, x_local: Tensor, num_channels: int, chunk: int, group=None,
                      out: Tensor | None = None) -> Tensor:
    """Filter this rank's ``[C_r, T]`` block in time chunks and all-gather every chunk while the
    next one is being filtered (the gather moves P x the kernel's bytes over NVLink,
    so it is hidden behind -- or rather hides -- the compute instead of following it).

    ``module`` must be a state-carrying SOS filter (IIR / Biquad / FusedSOSCascade): consecutive
    calls on consecutive chunks are then identical to one call (reference filter/iir.py:135-144).
    Chunk i is gathered on a side stream into a contiguous ``[P*C_r, n]`` staging buffer and
    copied into ``out[:, t0:t1]`` there; the compute stream only waits at the end.
    """
    if not _carries_state(module):
        raise TypeError(f"{type(module).__name__} does not carry state across calls; chunked gather needs an SOS filter")
    if chunk <= 0:
        raise ValueError(f"chunk must be positive, got {chunk}")
    world, _ = _world(group)
    C_r, T = x_local.shape
    if out is None:
        out = torch.empty((num_channels, T), dtype=x_local.dtype, device=x_local.device)
    if world == 1 or num_channels % world != 0 or num_channels != world * C_r:
        return all_gather_channels(module(x_local), num_channels, group, out)
    cuda = x_local.is_cuda
    if cuda:
        idx = x_local.device.index if x_local.device.index is not None else torch.cuda.current_device()
        comm = _comm_streams.get(idx)
        if comm is None:
            comm = _comm_streams[idx] = torch.cuda.Stream(device=x_local.device)
        compute = torch.cuda.current_stream(x_local.device)
        comm.wait_stream(compute)  # `out` and the staging buffers are allocated on the compute stream
    stage = [torch.empty(num_channels * min(chunk, T), dtype=x_local.dtype, device=x_local.device) for _ in range(2)]
    for i, t0 in enumerate(range(0, T, chunk)):
        n = min(chunk, T - t0)
        y = module(x_local[:, t0 : t0 + n])
        if not y.is_contiguous():
            y = y.contiguous()
        buf = stage[i % 2][: num_channels * n].view(num_channels, n)
        if cuda:
            comm.wait_stream(compute)
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(buf, y, group=group)
                out[:, t0 : t0 + n].copy_(buf)
            y.record_stream(comm)
        else:
            dist.all_gather_into_tensor(buf, y, group=group)
            out[:, t0 : t0 + n].copy_(buf)
    if cuda:
        compute.wait_stream(comm)
        for b in stage:
            b.record_stream(comm)
    return out


class ChannelSharded(nn.Module):
    """Run ``module`` on this rank's channel block of a ``[C, T]`` input.

    ``gather=False`` (default) returns the local block -- no collective on the data path.
    ``gather=True`` all-gathers the blocks so every rank returns the full ``[C, T]``; with
    ``gather_chunk`` (samples) and a state-carrying SOS filter the gather of each time chunk
    overlaps the filtering of the next one (``filter_and_gather``).
    The wrapped filter's state is rank-local (``[K, C/P, 2]``), coefficients are replicated.
    """

    def __init__(self, module: nn.Module, group=None, gather: bool = False, input_is_sharded: bool = False,
                 gather_chunk: int | None = None) -> None:
        super().__init__()
        self.module = module
        self.group = group
        self.gather = gather
        self.input_is_sharded = input_is_sharded
        self.gather_chunk = gather_chunk

    def forward(self, x: Tensor, num_channels: int | None = None) -> Tensor:
        if self.input_is_sharded:
            local = x
            if self.gather and num_channels is None:
                raise ValueError("num_channels is required to gather a pre-sharded input")
            total = num_channels
        else:
            total = x.shape[-2]
            local = shard_channels(x, self.group)
        if self.gather and self.gather_chunk and local.ndim == 2 and _carries_state(self.module):
            return filter_and_gather(self.module, local, int(total), self.gather_chunk, self.group)
        y = self.module(local)
        if not self.gather:
            return y
        return all_gather_channels(y, int(total), self.group)
