"""Secondary benches for BASELINE configs 3-5 on one GPU (the headline metric is bench.py).

Prints one JSON object per config: Gsamples/s (or lane-samples/s), algorithmic GB/s, fraction of the measured
HBM peak, and a parity spot-check against the CPU oracle."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy.signal as sps
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchfx_b200 as fx  # noqa: E402
from oracle import oracle  # noqa: E402
from torchfx_b200 import _native, _ops  # noqa: E402

FS = 48000
SECONDS = float(os.environ.get("CFG_SECONDS", "60"))
DEV = torch.device("cuda:0")
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _native.kernel_launches()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (_native.kernel_launches() - l0) // reps


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def cfg4():
    C, T = 2048, int(SECONDS * FS)
    x = torch.empty((C, T), device=DEV).normal_(0, 0.1)
    mk = lambda: [fx.filter.LoButterworth(5000, order=4, fs=FS), fx.filter.ParametricEQ(1000, q=2.0, gain=3.0, fs=FS),
                  fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=FS)]
    chain = mk()
    out = {}
    for prec in ("auto", "f32", "f64"):
        _ops.set_default_precision(prec)
        ms, nl = timed(lambda: (fx.Wave(x, FS, device=DEV) | chain[0] | chain[1] | chain[2]).ys)
        out[f"fused_{prec}"] = {"ms": round(ms, 3), "Gsamples_s": round(C * T / ms / 1e6, 1), "GBps": round(8 * C * T / ms / 1e6, 1),
                                "frac_hbm": round(8 * C * T / ms / 1e6 / PEAK, 3), "launches": nl}
    _ops.set_default_precision("auto")
    sep = mk()
    def unfused():
        y = x
        for f in sep:
            f.reset_state()
            f.fs = FS
            y = f(y)
        return y
    ms, nl = timed(unfused)
    out["unfused_auto"] = {"ms": round(ms, 3), "Gsamples_s": round(C * T / ms / 1e6, 1), "launches": nl}
    y = (fx.Wave(x[:4, : 1 << 18], FS, device=DEV) | mk()[0] | mk()[1] | mk()[2]).ys
    fs_ = mk()
    for f in fs_:
        f.compute_coefficients()
    sos = np.vstack([f._sos.numpy() for f in fs_])
    want, _, _ = oracle.sos_cascade(x[:4, : 1 << 18].cpu().numpy(), sos)
    out["parity_rel_err"] = rel(y.cpu().numpy(), want)
    return {"config": f"cfg4: fused LoButterworth|ParametricEQ|HiShelving, {C} ch x {SECONDS:g} s", **out}


def cfg5():
    out = {}
    C, T, N = 256, int(SECONDS * FS), 32
    x = torch.empty((C, T), device=DEV).normal_(0, 0.1)
    bank = fx.filter.LogFilterBank(n_bands=N, f_min=20.0, f_max=20000.0, q=1.414, fs=FS)
    ms, nl = timed(lambda: (bank.reset_state(), bank(x))[1], reps=3)
    lanes = N * C * T
    bytes_alg = 4 * lanes * (1 + 1 / N)
    out["stack"] = {"ms": round(ms, 3), "G_lane_samples_s": round(lanes / ms / 1e6, 1), "GBps": round(bytes_alg / ms / 1e6, 1),
                    "frac_hbm": round(bytes_alg / ms / 1e6 / PEAK, 3), "launches": nl, "shape": [N, C, T]}
    bank.reset_state()
    y = bank(x[:2, : 1 << 17])
    bank.compute_coefficients()
    want = oracle.filterbank_stack(x[:2, : 1 << 17].cpu().numpy(), np.stack([f._sos.numpy() for f in bank.filters]))
    out["stack"]["parity_rel_err"] = max(rel(y[b].cpu().numpy(), want[b]) for b in range(N))
    del y, x
    torch.cuda.empty_cache()
    C = 1024
    x = torch.empty((C, T), device=DEV).normal_(0, 0.1)
    fl = [fx.filter.BiquadBPF(200.0 * 1.7 ** i, 1.414, FS) for i in range(8)]
    comb = fx.filter._base.ParallelFilterCombination(*fl)
    def run_sum():
        for f in fl:
            f.reset_state()
        return comb(x)
    ms, nl = timed(run_sum, reps=3)
    out["sum"] = {"ms": round(ms, 3), "Gsamples_s": round(C * T / ms / 1e6, 1), "G_lane_samples_s": round(8 * C * T / ms / 1e6, 1),
                  "GBps": round(8 * C * T / ms / 1e6, 1), "frac_hbm": round(8 * C * T / ms / 1e6 / PEAK, 3), "launches": nl}
    for f in fl:
        f.reset_state()
    y = comb(x[:2, : 1 << 17])
    want = oracle.filterbank_sum(x[:2, : 1 << 17].cpu().numpy(), np.stack([f._sos.numpy() for f in fl]))
    out["sum"]["parity_rel_err"] = rel(y.cpu().numpy(), want)
    return {"config": f"cfg5: LogFilterBank(32) x 256 ch (stack) and 8 BiquadBPF + over 1024 ch (sum), {SECONDS:g} s", **out}


def cfg3():
    C, T, K = 256, int(SECONDS * FS), 65536
    rng = np.random.default_rng(7)
    ir = rng.standard_normal(K) * np.exp(-np.arange(K) / 8000.0)
    ir = (ir / np.sqrt((ir ** 2).sum())).astype(np.float32)
    x = torch.empty((C, T), device=DEV).normal_(0, 0.1)
    f = fx.filter.FIR(ir)
    ms, nl = timed(lambda: f(x), reps=3)
    out = {"ms": round(ms, 3), "Gsamples_s": round(C * T / ms / 1e6, 2), "GBps_algorithmic": round(8 * C * T / ms / 1e6, 1),
           "frac_hbm": round(8 * C * T / ms / 1e6 / PEAK, 4), "launches": nl,
           "flops_per_sample_est": 250, "TFLOPs_est": round(250 * C * T / ms / 1e9, 2)}
    y = f(x[:3, :150000])
    want = oracle.fir_causal(x[:3, :150000].cpu().numpy(), ir)
    out["parity_rel_err"] = rel(y.cpu().numpy(), want)
    # short kernels for context
    for k in (64, 1024):
        b = (rng.standard_normal(k) / np.sqrt(k)).astype(np.float32)
        fk = fx.filter.FIR(b)
        ms, _ = timed(lambda: fk(x), reps=3)
        out[f"taps_{k}"] = {"ms": round(ms, 3), "Gsamples_s": round(C * T / ms / 1e6, 1)}
    return {"config": f"cfg3: FIR overlap-save, 65536-tap IR, {C} ch x {SECONDS:g} s", **out}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg4", "cfg5", "cfg3"]
    for name in which:
        print(json.dumps(globals()[name]()), flush=True)
        torch.cuda.empty_cache()
