"""Developer check of the overlap-save FIR kernel on a GPU: a few shapes against the CPU oracle, then timings."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from torchfx_b200 import _native  # noqa: E402
from torchfx_b200.filter.fir import fir_causal  # noqa: E402

DEV = "cuda:0"


def check(C, T, K, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((C, T)).astype(np.float32)
    b = (rng.standard_normal(K) * np.exp(-np.arange(K) / (K / 5.0))).astype(np.float32)
    t0 = time.perf_counter()
    y = fir_causal(torch.from_numpy(x).to(DEV), torch.from_numpy(b), _native.TFX_FIR_OLS)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    want = oracle.fir_causal(x, b)
    err = np.abs(y.cpu().numpy() - want).max(axis=1) / np.abs(want).max(axis=1)
    print(f"C={C} T={T} K={K}: rel err max {err.max():.3e} (per channel {np.array2string(err[:6], precision=2)}), {dt * 1e3:.2f} ms", flush=True)
    return err.max()


def bench(C, T, K, reps=5):
    x = torch.empty((C, T), device=DEV).normal_(0, 0.1)
    b = torch.randn(K) * torch.exp(-torch.arange(K) / 8000.0)
    fir_causal(x, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fir_causal(x, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"bench C={C} T={T} K={K}: {ms:.3f} ms = {C * T / ms / 1e6:.1f} Gsamples/s", flush=True)


if __name__ == "__main__":
    worst = 0.0
    for C, T, K in [(2, 8192, 100), (2, 20000, 3000), (1, 50001, 9000), (3, 100000, 8192), (5, 300000, 65536), (4, 70000, 70000),
                    (17, 600000, 20000), (2, 1000, 300)]:
        worst = max(worst, check(C, T, K, seed=C + T))
    print("worst", worst)
    if worst < 1e-5 and "--bench" in sys.argv:
        bench(256, 2880000, 65536)
        bench(256, 2880000, 1024)
        bench(256, 2880000, 20000)
        bench(2, 2880000, 65536)
