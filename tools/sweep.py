"""Developer perf sweep (not a bench value): times the cascade kernel over shapes/precisions."""
from __future__ import annotations

import json
import os
import sys

import scipy.signal as sps
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchfx_b200 import _ops  # noqa: E402


def time_call(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    rows = []
    cases = []
    secs = float(os.environ.get("SWEEP_SECONDS", "60"))
    for C in (1024, 128, 2048):
        for K in (1, 2, 4, 8):
            for prec in ("f32", "f64"):
                cases.append((C, int(secs * 48000 * 1024 / C), K, prec, False))
    cases.append((1024, int(secs * 48000), 4, "f32", True))
    cases.append((1000, 2880001, 4, "f32", False))  # misaligned rows -> element-wise path
    for C, T, K, prec, no_split in cases:
        x = torch.empty((C, T), dtype=torch.float32, device=dev).normal_(0, 0.1)
        sos = torch.from_numpy(sps.butter(2 * K, 5000 / 24000, output="sos")).contiguous()
        ms = time_call(lambda: _ops.sos_cascade_(x, sos, None, None, out=x, precision=prec, no_split=no_split))
        gbs = 8 * C * T / ms / 1e6
        rows.append({"C": C, "T": T, "K": K, "prec": prec, "no_split": no_split, "ms": round(ms, 3), "GBps": round(gbs, 1),
                     "Gsamples": round(C * T / ms / 1e6, 1)})
        print(rows[-1], flush=True)
        del x
    # plain copy for reference
    a = torch.empty(1024 * 2880000, dtype=torch.float32, device=dev).normal_()
    b = torch.empty_like(a)
    ms = time_call(lambda: b.copy_(a))
    print({"copy_GBps": round(8 * a.numel() / ms / 1e6, 1)})
    json.dump(rows, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "sweep.json"), "w"))


if __name__ == "__main__":
    main()
