"""Developer tool: run the FIR path a few times on one shape (target for ncu captures)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
C = int(os.environ.get("OS_C", 256)); T = int(os.environ.get("OS_T", 2880000)); K = int(os.environ.get("OS_K", 65536))
reps = int(os.environ.get("OS_REPS", 2))
rng = np.random.default_rng(7)
ir = rng.standard_normal(K) * np.exp(-np.arange(K) / 8000.0)
ir = (ir / np.sqrt((ir ** 2).sum())).astype(np.float32)
x = torch.empty((C, T), dtype=torch.float32, device="cuda").normal_(0, 0.1)
f = fx.filter.FIR(ir)
for _ in range(reps):
    y = f(x)
torch.cuda.synchronize()
print("done", tuple(y.shape))
