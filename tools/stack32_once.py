"""Developer tool: one STACK LogFilterBank(32) call on the 8-GPU shard shape of config 5 (target for ncu launch lists)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
C = int(os.environ.get("OS_C", 32)); T = int(os.environ.get("OS_T", 2880000))
x = torch.empty((C, T), device="cuda").normal_(0, 0.1)
bank = fx.filter.LogFilterBank(n_bands=32, f_min=20.0, f_max=20000.0, q=1.414, fs=48000)
for _ in range(3):
    bank.reset_state()
    y = None; y = bank(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    bank.reset_state(); y = None; y = bank(x)
e1.record(); torch.cuda.synchronize()
print("ms per call", e0.elapsed_time(e1) / 5, tuple(y.shape))
