"""Developer tool: float64-recurrence cascade (float32 I/O) through the TMA kernel (default for f64) and the
LDGSTS tile kernel (no_tma), plus K sweeps."""
import os, sys, json
import scipy.signal as sps, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchfx_b200 import _ops
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
C, T = 1024, 2880000
x = torch.empty((C, T), dtype=torch.float32, device="cuda").normal_(0, 0.1)
out = {}
PREC = os.environ.get("OS_PREC", "f64")
for K in [int(k) for k in os.environ.get("OS_KS", "1,2,4,8").split(",")]:
    sos = torch.from_numpy(sps.butter(2 * K, 5000 / 24000, output="sos")).contiguous()
    for name, kw in (("tma", {"force_tma": True}), ("tile", {"no_tma": True})):
        ms = t(lambda: _ops.sos_cascade_(x, sos, None, None, out=x, precision=PREC, **kw))
        out[f"{PREC}_K{K}_{name}"] = [round(ms, 3), round(C * T / ms / 1e6, 1)]
print(json.dumps(out))
