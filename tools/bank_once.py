"""Developer tool: run the STACK filterbank a few times on one shape (target for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
from torchfx_b200 import _ops
C = int(os.environ.get("OS_C", 256)); T = int(os.environ.get("OS_T", 480000)); N = int(os.environ.get("OS_N", 32))
prec = os.environ.get("OS_PREC", "f32"); reps = int(os.environ.get("OS_REPS", 2))
_ops.set_default_precision(prec)
x = torch.empty((C, T), dtype=torch.float32, device="cuda").normal_(0, 0.1)
bank = fx.filter.LogFilterBank(n_bands=N, f_min=20.0, f_max=20000.0, q=1.414, fs=48000)
for _ in range(reps):
    bank.reset_state()
    y = bank(x)
torch.cuda.synchronize()
print("done", tuple(y.shape))
