"""Developer tool: run the cascade a few times on one shape (target for ncu captures)."""
import os, sys
import scipy.signal as sps, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchfx_b200 import _ops
C = int(os.environ.get("OS_C", 1024)); T = int(os.environ.get("OS_T", 1440000)); K = int(os.environ.get("OS_K", 4))
prec = os.environ.get("OS_PREC", "f32"); reps = int(os.environ.get("OS_REPS", 3))
x = torch.empty((C, T), dtype=torch.float32, device="cuda").normal_(0, 0.1)
sos = torch.from_numpy(sps.butter(2 * K, 5000 / 24000, output="sos")).contiguous()
for _ in range(reps):
    _ops.sos_cascade_(x, sos, None, None, out=x, precision=prec)
torch.cuda.synchronize()
print("done")
