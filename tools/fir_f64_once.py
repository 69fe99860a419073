"""Developer tool: the float64 direct-form FIR (tfx_fir_f64) on 64 ch x 10 s at a few tap counts."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchfx_b200.filter.fir import fir_causal
x = torch.empty((64, 480000), device="cuda", dtype=torch.float64).normal_(0, 0.1)
for K in (8, 64, 512, 513, 4096):
    b = torch.randn(K, dtype=torch.float64)
    fir_causal(x, b); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fir_causal(x, b)
    e1.record(); torch.cuda.synchronize()
    print(f"f64 direct FIR 64 ch x 480000, K={K}: {e0.elapsed_time(e1) / 3:.3f} ms", flush=True)
