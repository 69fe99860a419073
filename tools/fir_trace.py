"""Developer tool: per-item trace of the persistent overlap-save FIR kernel (TFX_FIR_TRACE=1).

Prints, per item type (forward FFT / multiply-accumulate / inverse FFT), how many items ran, how long they ran and
how long they waited on their dependencies, plus the kernel span and the busy fraction of the SMs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torchfx_b200 import _native as N  # noqa: E402
from torchfx_b200.filter.fir import fir_causal  # noqa: E402

DEV = torch.device("cuda:0")
NCTA = 296 if os.environ.get("TFX_FIR_N") == "8192" else 148  # resident CTAs of the persistent kernel


def main(C=256, T=2880000, K=65536):
    lib = N.load()
    os.environ.pop("TFX_FIR_TRACE", None)
    off = lib.tfx_fir_workspace_bytes(C, T, K, N.TFX_FIR_OLS)
    os.environ["TFX_FIR_TRACE"] = "1"
    x = torch.empty((C, T), device=DEV).normal_(0, 0.1)
    b = torch.randn(K) * torch.exp(-torch.arange(K) / 8000.0)
    for _ in range(2):
        fir_causal(x, b, N.TFX_FIR_OLS)
    torch.cuda.synchronize()
    ws = N._workspaces[(0, torch.cuda.current_stream(DEV).cuda_stream)]
    ph = ws[off + (1 << 20) * 16: off + (1 << 20) * 16 + 256].view(torch.int64).cpu().numpy()
    tr = ws[off:off + (1 << 20) * 16].view(torch.int32).reshape(-1, 4).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    used = tr[:, 2] > 0
    tr = tr[used | (tr[:, 1] > 0)]
    typ, smid = tr[:, 0] >> 24, (tr[:, 0] >> 8) & 0xFFFF
    wait, run, start = tr[:, 1], tr[:, 2], tr[:, 3]
    t0 = start.min()
    rel = (start - t0) & 0xFFFFFFFF
    span = (rel + wait + run).max()
    print(f"C={C} T={T} K={K}: {len(tr)} traced items, kernel span {span / 1e6:.3f} ms, SMs seen {len(np.unique(smid))}")
    for t, name in enumerate(("forward", "mac", "inverse")):
        m = (typ == t) & (run > 2000)
        if not m.any():
            continue
        print(f"  {name:8s} n={m.sum():6d}  run mean {run[m].mean() / 1e3:7.2f} us  median {np.median(run[m]) / 1e3:7.2f}  p95 {np.percentile(run[m], 95) / 1e3:7.2f}"
              f"  | wait mean {wait[m].mean() / 1e3:7.2f} us  p95 {np.percentile(wait[m], 95) / 1e3:7.2f}  | sum run {run[m].sum() / 1e6 / NCTA:7.3f} ms/SM  sum wait {wait[m].sum() / 1e6 / NCTA:7.3f} ms/SM")
    idle = (typ >= 0) & (run <= 2000)
    print(f"  skipped/empty items: {idle.sum()}, their waits sum {wait[idle].sum() / 1e6 / NCTA:.3f} ms/SM")
    names = ["mac wait+barrier", "mac issue", "mac compute+store", "fwd load+P1+barrier", "fwd P2+P3", "fwd P4+store", "inv load+P4", "inv P3+P2",
             "inv barrier", "inv P1+store", "fwd signal", "mac signal", "inv signal"]
    n_by = {"mac": (typ == 1).sum(), "fwd": (typ == 0).sum(), "inv": (typ == 2).sum()}
    print("  phase clocks (thread 0, cycles per item of that type):")
    for i, nm in enumerate(names):
        print(f"    {nm:22s} {ph[i] / max(n_by[nm[:3]], 1):9.0f}")
    busy = run.sum() / (NCTA * span)
    print(f"  busy fraction (run / (CTAs x span)) = {busy:.3f}, waiting fraction = {wait.sum() / (NCTA * span):.3f}")


if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:]]
    main(*args)
