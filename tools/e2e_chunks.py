"""Developer tool: the host-buffer streaming entry (tfx_sos_cascade_host_f32, what bench.py times as `e2e`) against the chunk
length, on config 2's 1024 ch x 60 s slice in pinned memory."""
import os, sys, time
import numpy as np, scipy.signal as sps, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchfx_b200 import _native
lib = _native.load()
C, Te = 1024, 60 * 48000
sos = torch.from_numpy(sps.butter(8, 5000 / 24000, output="sos")).contiguous()
xh = torch.empty((C, Te), dtype=torch.float32, pin_memory=True); xh.normal_(0.0, 0.1)
yh = torch.empty((C, Te), dtype=torch.float32, pin_memory=True)
for chunk in [int(a) for a in sys.argv[1:]] or [0, 16384, 32768, 131072, 262144]:
    def step():
        _native.check(lib.tfx_sos_cascade_host_f32(xh.data_ptr(), yh.data_ptr(), C, Te, Te, Te, sos.data_ptr(), 4, None, None, 0, chunk, 0))
    step()
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    dt = (time.perf_counter() - t0) / 3
    print(f"chunk_T={chunk or 'default (65536)'}: {dt * 1e3:.1f} ms per step = {C * Te / dt / 1e9:.2f} Gsamples/s ({8 * C * Te / dt / 1e9 / 2:.1f} GB/s each way)", flush=True)
