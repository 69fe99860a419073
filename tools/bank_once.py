"""Developer tool: run a filterbank a few times on one shape (target for ncu captures).
OS_MODE=stack: LogFilterBank(OS_N) over OS_C channels; OS_MODE=sum: OS_N BiquadBPF summed (`+`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
from torchfx_b200 import _ops
mode = os.environ.get("OS_MODE", "stack")
C = int(os.environ.get("OS_C", 256 if mode == "stack" else 1024)); T = int(os.environ.get("OS_T", 480000))
N = int(os.environ.get("OS_N", 32 if mode == "stack" else 8))
prec = os.environ.get("OS_PREC", "f32"); reps = int(os.environ.get("OS_REPS", 2))
_ops.set_default_precision(prec)
x = torch.empty((C, T), dtype=torch.float32, device="cuda").normal_(0, 0.1)
if mode == "stack":
    bank = fx.filter.LogFilterBank(n_bands=N, f_min=20.0, f_max=20000.0, q=1.414, fs=48000)
    run = lambda: (bank.reset_state(), bank(x))[1]
else:
    fl = ([fx.filter.BiquadBPF(200.0 * 1.7 ** i, 1.414, 48000) for i in range(N)] if N <= 8 else
          [fx.filter.BiquadBPF(20.0 * (1000.0 ** (i / (N - 1.0))), 1.414, 48000) for i in range(N)])  # config 5 read literally
    comb = fx.filter._base.ParallelFilterCombination(*fl)
    def run():
        for f in fl:
            f.reset_state()
        return comb(x)
for _ in range(reps):
    y = run()
    del y
torch.cuda.synchronize()
print("done")
