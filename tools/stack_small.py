"""Developer tool: STACK LogFilterBank(32) on shapes with few channel groups (the 8-GPU / 2-GPU shards of config 5):
default dispatch (bank_stack with the bands split over CTAs) against the band-per-lane kernel (TFX_NO_TILE)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
from torchfx_b200 import _native
from torchfx_b200.filter._sosbank import SosBank
def t(fn, reps=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
for C, T in ((32, 2880000), (64, 2880000), (128, 480000), (128, 2880000), (256, 480000)):
    x = torch.empty((C, T), device="cuda").normal_(0, 0.1)
    for name, flags in (("default", 0), ("band_per_lane", _native.TFX_NO_TILE), ("forced_tile", _native.TFX_FORCE_TILE)):
        mk = [fx.filter.BiquadBPF(20.0 * (1000.0 ** (i / 31.0)), 1.414, 48000) for i in range(32)]
        bank = SosBank(mk, mode="stack"); bank.flags = flags
        def run():
            for f in mk: f.reset_state()
            return bank(x)
        ms = t(run)
        out[f"{C}x{T}_{name}"] = [round(ms, 3), round(32 * C * T / ms / 1e6)]
    del x
print(json.dumps(out))
