"""Developer tool: config 4 (fused LoButterworth | ParametricEQ | HiShelving, TFX_PREC_AUTO) on the channel shard one GPU
gets at N GPUs (2048 / N channels x 60 s), against 1 / N of the full job."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
from torchfx_b200 import _ops
T = 2880000
chain = [fx.filter.LoButterworth(5000, order=4, fs=48000), fx.filter.ParametricEQ(1000, q=2.0, gain=3.0, fs=48000),
         fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=48000)]
for f in chain:
    f.compute_coefficients()
sos = torch.from_numpy(np.vstack([f._sos.numpy() for f in chain])).contiguous()
for C in (2048, 1024, 512, 256):
    x = torch.empty((C, T), device="cuda").normal_(0, 0.1)
    y = torch.empty_like(x)
    fn = lambda: _ops.sos_cascade_(x, sos, None, None, out=y)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"TFX_WARM_DIV={os.environ.get('TFX_WARM_DIV', 'default')} C={C}: {ms:.3f} ms/call, {C * T / ms / 1e6:.1f} Gsamples/s = {8 * C * T / ms / 1e6 / 6550.1:.3f} of HBM", flush=True)
    del x, y
    torch.cuda.empty_cache()
