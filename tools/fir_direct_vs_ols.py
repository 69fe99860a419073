import sys, os, torch
sys.path.insert(0, "/root/repo")
from torchfx_b200 import _native
from torchfx_b200.filter.fir import fir_causal
def bench(C, T, K, algo, reps=5):
    x = torch.empty((C, T), device="cuda").normal_(0, 0.1)
    b = torch.randn(K)
    fir_causal(x, b, algo); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fir_causal(x, b, algo)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for C, T in ((256, 2880000), (2, 48000), (8, 480000), (64, 480000)):
    for K in (8, 16, 32, 48, 64, 96):
        d = bench(C, T, K, _native.TFX_FIR_DIRECT); o = bench(C, T, K, _native.TFX_FIR_OLS)
        print(f"C={C} T={T} K={K}: direct {d:.3f} ms, overlap-save {o:.3f} ms", flush=True)
