"""Developer tool: time the 32-band SUM bank (config 5 read literally) on one shape; TFX_BS_SUM_WARPS caps the warps per CTA."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
from torchfx_b200 import _ops
C, T, N = int(os.environ.get("OS_C", 256)), int(os.environ.get("OS_T", 2880000)), int(os.environ.get("OS_N", 32))
x = torch.empty((C, T), device="cuda").normal_(0, 0.1)
for prec in ("f32", "auto"):
    _ops.set_default_precision(prec)
    fl = [fx.filter.BiquadBPF(20.0 * (1000.0 ** (i / (N - 1.0))), 1.414, 48000) for i in range(N)]
    comb = fx.filter._base.ParallelFilterCombination(*fl)
    def run():
        for f in fl:
            f.reset_state()
        return comb(x)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"sum{N} x {C} ch x {T}: prec={prec} {ms:.3f} ms = {C * T / ms / 1e6:.1f} Gsamples/s, TFX_BS_SUM_WARPS={os.environ.get('TFX_BS_SUM_WARPS', '-')}", flush=True)
