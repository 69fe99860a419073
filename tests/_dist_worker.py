"""Worker for tests/test_dist.py: run under torch.distributed.run with the gloo backend."""
import os
import sys

import numpy as np
import scipy.signal as sps
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx  # noqa: E402
from torchfx_b200.dist import ChannelSharded, all_gather_channels, filter_and_gather, shard_bounds, shard_channels  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(123)  # same input on every rank
    for C in (8, 7):  # even split and ragged split
        x = torch.randn(C, 6000, dtype=torch.float64)
        f = fx.filter.LoButterworth(3000, order=6, fs=48000)
        ref = sps.sosfilt(sps.butter(6, 3000 / 24000, output="sos"), x.numpy(), axis=-1)
        # no collective on the data path: each rank filters its own block
        local = ChannelSharded(f)(x)
        lo, hi = shard_bounds(C, world, rank)
        assert local.shape == (hi - lo, 6000)
        np.testing.assert_allclose(local.numpy(), ref[lo:hi], atol=1e-10)
        assert f._state_x.shape == (3, hi - lo, 2)  # state is rank-local
        # optional reassembly: all-gather of the channel blocks
        full = all_gather_channels(local, C)
        np.testing.assert_allclose(full.numpy(), ref, atol=1e-10)
        # streaming: two chunks with rank-local state carry, gathered
        g = ChannelSharded(fx.filter.LoButterworth(3000, order=6, fs=48000), gather=True)
        y = torch.cat([g(x[:, :2500]), g(x[:, 2500:])], dim=1)
        np.testing.assert_allclose(y.numpy(), ref, atol=1e-10)
        # chunked gather (each time chunk gathered as soon as it is filtered; overlapped on GPUs)
        gc = ChannelSharded(fx.filter.LoButterworth(3000, order=6, fs=48000), gather=True, gather_chunk=1024)
        np.testing.assert_allclose(gc(x).numpy(), ref, atol=1e-10)
    try:
        filter_and_gather(fx.filter.FIR([1.0, 0.5]), torch.randn(2, 100), 4, 32)
    except TypeError as e:
        assert "does not carry state" in str(e)
    else:
        raise AssertionError("filter_and_gather accepted a stateless module")
    # filterbank lanes: shard channels, gather per band
    x = torch.randn(4, 3000)
    bank = fx.filter.LogFilterBank(n_bands=4, f_min=200.0, f_max=4000.0, fs=48000)
    yb = bank(shard_channels(x))  # [bands, C/P, T]
    gathered = torch.stack([all_gather_channels(yb[b].contiguous(), 4) for b in range(4)])
    whole = fx.filter.LogFilterBank(n_bands=4, f_min=200.0, f_max=4000.0, fs=48000)(x)
    torch.testing.assert_close(gathered, whole, atol=1e-6, rtol=0)
    # the union of shards is exactly the channel range
    cover = sorted(c for r in range(world) for c in range(*shard_bounds(7, world, r)))
    assert cover == list(range(7))
    dist.barrier()
    if rank == 0:
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
