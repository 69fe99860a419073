"""Developer tool: the cascade on the channel shard ONE GPU gets at N GPUs (config 2: 1024 / N channels x 10 min, in place),
against the ideal 1 / N of the full job.  TFX_WARM_DIV overrides the planner's segment-length rule (sos_plan.cpp)."""
import os, sys
import scipy.signal as sps, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchfx_b200 import _ops
T = 28800000
sos = torch.from_numpy(sps.butter(8, 5000 / 24000, output="sos")).contiguous()
for C in (1024, 512, 256, 128):
    x = torch.empty((C, T), device="cuda").normal_(0, 0.1)
    fn = lambda: _ops.sos_cascade_(x, sos, None, None, out=x)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"TFX_WARM_DIV={os.environ.get('TFX_WARM_DIV', 'default')} C={C}: {ms:.3f} ms/step, {C * T / ms / 1e6:.1f} Gsamples/s, x{1024 // C} = {1024 * T / ms / 1e6 * 1:.0f} job-equivalent", flush=True)
    del x
    torch.cuda.empty_cache()
