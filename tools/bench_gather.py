"""Kernel-only vs kernel + all-gather for the sharded configs, run under torchrun."""
import json, os, sys
import scipy.signal as sps, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchfx_b200 as fx
from torchfx_b200 import _ops
from torchfx_b200.dist import all_gather_channels, filter_and_gather, shard_bounds

def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    FS, SEC = 48000, float(os.environ.get("GATHER_SECONDS", "10"))
    T = int(SEC * FS)
    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    out = {"n_gpus": world, "seconds": SEC}
    # cascade, 1024 channels total (cfg2 shape at reduced length)
    C = 1024; lo, hi = shard_bounds(C, world, rank)
    x = torch.empty((hi - lo, T), device=dev).normal_(0, 0.1)
    sos = torch.from_numpy(sps.butter(8, 5000 / 24000, output="sos")).contiguous()
    full = torch.empty((C, T), device=dev)
    y = torch.empty_like(x)
    k_ms = timed(lambda: _ops.sos_cascade_(x, sos, None, None, out=y))
    g_ms = timed(lambda: all_gather_channels(_ops.sos_cascade_(x, sos, None, None, out=y), C, out=full))
    gathered = 4.0 * C * T * (world - 1) / world  # bytes received per rank
    # overlapped: 1 s time chunks, chunk i gathered on a side stream while chunk i+1 is filtered
    fused = fx.filter.FusedSOSCascade(fx.filter.LoButterworth(5000, order=8, fs=FS))
    def run_overlapped():
        fused.reset_state(); return filter_and_gather(fused, x, C, FS, out=full)
    o_ms = timed(run_overlapped)
    ref_full = all_gather_channels(_ops.sos_cascade_(x, sos, None, None), C)
    ov_err = float((run_overlapped() - ref_full).abs().max() / ref_full.abs().max())
    del ref_full
    out["cascade"] = {"kernel_ms": round(k_ms, 3), "kernel_plus_gather_ms": round(g_ms, 3), "Gsamples_s_kernel": round(C * T / k_ms / 1e6, 1),
                      "Gsamples_s_with_gather": round(C * T / g_ms / 1e6, 1), "gather_GBps_per_rank": round(gathered / max(g_ms - k_ms, 1e-9) / 1e6, 1),
                      "overlapped_chunked_ms": round(o_ms, 3), "Gsamples_s_overlapped": round(C * T / o_ms / 1e6, 1),
                      "overlapped_link_GBps_per_rank": round(gathered / o_ms / 1e6, 1), "overlapped_vs_unchunked_rel_diff": ov_err}
    del x, y, full
    # filterbank stack: 32 bands x 256 channels total, gathered per band plane
    C, N = 256, 32; lo, hi = shard_bounds(C, world, rank)
    x = torch.empty((hi - lo, T), device=dev).normal_(0, 0.1)
    bank = fx.filter.LogFilterBank(n_bands=N, f_min=20.0, f_max=20000.0, q=1.414, fs=FS)
    full = torch.empty((N, C, T), device=dev)
    def run_bank():
        bank.reset_state(); return bank(x)
    def run_bank_gather():
        yb = run_bank()
        for b in range(N): all_gather_channels(yb[b], C, out=full[b])
    k_ms = timed(run_bank, reps=3)
    g_ms = timed(run_bank_gather, reps=3)
    out["filterbank_stack"] = {"kernel_ms": round(k_ms, 3), "kernel_plus_gather_ms": round(g_ms, 3),
                               "G_lane_samples_s_kernel": round(N * C * T / k_ms / 1e6, 1), "G_lane_samples_s_with_gather": round(N * C * T / g_ms / 1e6, 1),
                               "gather_GBps_per_rank": round(4.0 * N * C * T * (world - 1) / world / max(g_ms - k_ms, 1e-9) / 1e6, 1)}
    if rank == 0: print(json.dumps(out), flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
