"""Developer tool: a few small overlap-save FIR calls (target of compute-sanitizer memcheck / racecheck runs)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from torchfx_b200 import _native
from torchfx_b200.filter.fir import fir_causal
for C, T, K in [(3, 40000, 9000), (2, 20000, 300), (4, 70000, 20000)]:
    rng = np.random.default_rng(C + T)
    x = rng.standard_normal((C, T)).astype(np.float32)
    b = (rng.standard_normal(K) * np.exp(-np.arange(K) / (K / 5.0))).astype(np.float32)
    y = fir_causal(torch.from_numpy(x).cuda(), torch.from_numpy(b), _native.TFX_FIR_OLS)
    torch.cuda.synchronize()
    want = oracle.fir_causal(x, b)
    print(C, T, K, float(np.abs(y.cpu().numpy() - want).max() / np.abs(want).max()), flush=True)
