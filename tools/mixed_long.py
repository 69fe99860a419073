"""Developer tool: cascades of 5..8 sections with float32-hostile sections, in place, 1024 ch x 60 s: TFX_PREC_AUTO (mixed-precision
tile kernel, sos_tile_mixed.cu) against the all-float32 and all-float64 recurrences."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
from torchfx_b200 import _ops
FS = 48000
C, T = 1024, int(os.environ.get("OS_T", 2880000))
x = torch.empty((C, T), device="cuda").normal_(0, 0.1)
lo = fx.filter.LoButterworth(6000, order=6, fs=FS)
shelf = fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=FS)
hp = fx.filter.HiButterworth(20, order=2, fs=FS)
notch = fx.filter.Notch(60, q=30.0, fs=FS)
eq = fx.filter.ParametricEQ(9000, q=1.0, gain=-2.0, fs=FS)
for name, chain in {"K=6, 20 Hz high-pass first": [hp, lo, shelf, eq], "K=6, high-pass in the middle": [lo, hp, shelf, eq],
                    "K=7, high-pass + notch": [lo, hp, shelf, notch, eq]}.items():
    for f in chain:
        f.compute_coefficients()
    sos = torch.from_numpy(np.vstack([f._sos.numpy() for f in chain])).contiguous()
    res = {}
    for prec in ("auto", "f32", "f64"):
        fn = lambda: _ops.sos_cascade_(x, sos, None, None, out=x, precision=prec)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[prec] = round(C * T / ms / 1e6, 1)
        x.normal_(0, 0.1)
    print(f"{name}: Gsamples/s {res}", flush=True)
