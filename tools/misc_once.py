"""Developer tool: small drivers for the kernels that had no ncu capture in round 1 (target of tools/ncu_round2.sh).
OS_WHAT = delay | fir_direct | fir_f64 | bank_stream | mixed6"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
from torchfx_b200 import _ops
what = os.environ.get("OS_WHAT", "delay")
reps = int(os.environ.get("OS_REPS", 3))
FS = 48000
if what == "delay":
    x = torch.empty((1024, 60 * FS), device="cuda").normal_(0, 0.1)
    fn = lambda: fx.Reverb(delay=4410, decay=0.5, mix=0.5)(x)
elif what == "fir_direct":
    x = torch.empty((256, 60 * FS), device="cuda").normal_(0, 0.1)
    f = fx.filter.FIR(np.hanning(64).astype(np.float32), conv_mode="direct")  # (AUTO takes overlap-save above 56 taps at this size)
    fn = lambda: f(x)
elif what == "fir_f64":
    x = torch.empty((16, 10 * FS), device="cuda", dtype=torch.float64).normal_(0, 0.1)
    f = fx.filter.FIR(np.hanning(512).astype(np.float32))
    fn = lambda: f(x)
elif what == "bank_stream":
    x = torch.empty((4, 60 * FS), device="cuda").normal_(0, 0.1)
    bank = fx.filter.LogFilterBank(n_bands=32, f_min=20.0, f_max=20000.0, q=1.414, fs=FS)
    fn = lambda: (bank.reset_state(), bank(x))[1]
elif what == "mixed6":
    x = torch.empty((1024, 30 * FS), device="cuda").normal_(0, 0.1)
    chain = [fx.filter.LoButterworth(6000, order=6, fs=FS), fx.filter.HiButterworth(20, order=2, fs=FS),
             fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=FS), fx.filter.ParametricEQ(9000, q=1.0, gain=-2.0, fs=FS)]
    fn = lambda: (fx.Wave(x, FS, device="cuda") | chain[0] | chain[1] | chain[2] | chain[3]).ys
else:
    raise SystemExit(what)
for _ in range(reps):
    y = fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    y = fn()
e1.record(); torch.cuda.synchronize()
print(what, "ms per call", e0.elapsed_time(e1) / reps, tuple(y.shape), y.dtype)
