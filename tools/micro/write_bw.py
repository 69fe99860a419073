"""Developer probe: write-only, read-only and copy HBM bandwidth with torch's vectorised kernels
(what a store-dominated kernel such as the STACK filterbank can expect at best)."""
import torch
n = 12 * (1 << 30)  # 12 Gi floats = 48 GiB
a = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.fill_(1.0)); print(f"fill_  (write only): {4*n/ms/1e6:.0f} GB/s")
ms = t(lambda: torch.cuda.memset if False else a.zero_()); print(f"zero_  (write only): {4*n/ms/1e6:.0f} GB/s")
ms = t(lambda: a.sum()); print(f"sum    (read only) : {4*n/ms/1e6:.0f} GB/s")
b = torch.empty_like(a)
ms = t(lambda: b.copy_(a)); print(f"copy_  (read+write): {8*n/ms/1e6:.0f} GB/s")
# 1 read : 32 writes like the 32-band STACK bank: expand
c = torch.empty((1 << 28,), dtype=torch.float32, device="cuda")
d = torch.empty((32, 1 << 28), dtype=torch.float32, device="cuda")
del a, b
ms = t(lambda: d.copy_(c.unsqueeze(0).expand(32, -1))); print(f"expand copy 1->32  : {4*33*(1<<28)/ms/1e6:.0f} GB/s")
