"""Secondary comparison (SURVEY.md 8d last row): the reference's OWN CUDA path, compiled
unmodified for sm_100 (`make -C oracle refcuda` -> oracle/_ref/cuda/torchfx_ext.so), timed
on the same B200 next to this library, on a configuration the reference can run
(C*T < 2^31 and ~40 B/sample of float64 temporaries): 64 channels x 60 s, K = 4.

The reference call is reproduced with its glue (src/torchfx/_ops.py:119-176 and
filter/iir.py:174-176): f32 -> f64 cast, `sos_forward` (a Blelloch scan per section,
cuda/parallel_scan.cu:117-364), f64 -> f32 cast.  Test infrastructure, like oracle/.

    python tools/bench_ref_cuda.py            # prints one JSON object
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import json
import os
import sys

import numpy as np
import scipy.signal as sps
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, "oracle", "_ref", "cuda", "torchfx_ext.so")


def load_ref_cuda():
    loader = importlib.machinery.ExtensionFileLoader("torchfx_ext", SO)
    spec = importlib.util.spec_from_loader("torchfx_ext", loader, origin=SO)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def time_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.median(out))


def main():
    from torchfx_b200 import _ops

    if not os.path.exists(SO):
        print(json.dumps({"unavailable": f"{SO} missing (make -C oracle refcuda)"}))
        return
    ext = load_ref_cuda()
    dev = torch.device("cuda", 0)
    res = {}
    for C, seconds in ((64, 60), (256, 30), (8, 60)):
        T = seconds * 48000
        g = torch.Generator(device=dev).manual_seed(1234)
        x = 0.1 * torch.randn(C, T, device=dev, generator=g)
        sos_cpu = torch.as_tensor(sps.butter(8, 5000 / 24000, output="sos"), dtype=torch.float64)
        sos_dev = sos_cpu.to(dev)
        K = sos_cpu.shape[0]

        def ref_call():
            sx = torch.zeros(K, C, 2, device=dev, dtype=torch.float64)
            sy = torch.zeros(K, C, 2, device=dev, dtype=torch.float64)
            y64, _, _ = ext.sos_forward(x.to(torch.float64), sos_dev, sos_cpu, sx, sy)
            return y64.to(torch.float32)

        def ref_kernel_only(x64=None):
            sx = torch.zeros(K, C, 2, device=dev, dtype=torch.float64)
            sy = torch.zeros(K, C, 2, device=dev, dtype=torch.float64)
            return ext.sos_forward(x64, sos_dev, sos_cpu, sx, sy)[0]

        def ours():
            return _ops.sos_cascade_(x, sos_cpu, None, None)

        y_ref = ref_call()
        y_ours = ours()
        torch.cuda.synchronize()
        diff = float((y_ref - y_ours).abs().max() / y_ref.abs().max())
        t_ref = time_ms(ref_call)
        x64 = x.to(torch.float64)
        t_ref_k = time_ms(lambda: ref_kernel_only(x64))
        del x64
        t_ours = time_ms(ours, reps=20, warm=3)
        n = C * T
        res[f"{C}ch_x_{seconds}s"] = {
            "reference_cuda_ms": round(t_ref, 3),
            "reference_cuda_Gsamples_s": round(n / t_ref / 1e6, 2),
            "reference_cuda_f64_native_call_only_ms": round(t_ref_k, 3),
            "ours_ms": round(t_ours, 4),
            "ours_Gsamples_s": round(n / t_ours / 1e6, 2),
            "speedup": round(t_ref / t_ours, 1),
            "max_rel_diff_ours_vs_reference_cuda": diff,
        }
        del x, y_ref, y_ours
        torch.cuda.empty_cache()
    print(json.dumps({"reference_cuda_vs_ours_K4": res}))


if __name__ == "__main__":
    main()
