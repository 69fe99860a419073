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


def cpu_reference(kind: str):
    """Reference CPU path timed beside the GPU number, on a bounded slice:
    the unmodified reference extension (oracle/_ref) for the SOS paths, the numpy restatement of
    the reference's torch.fft overlap-save for FIR (the reference Python package cannot travel)."""
    import time

    from oracle import ref_loader

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ext = ref_loader.load_ref_ext() if ref_loader.have_ref_ext() else None

    def sos_call(x32, sos_np):
        sos_t = torch.from_numpy(np.ascontiguousarray(sos_np))
        if ext is None:
            return oracle.sos_cascade(x32.numpy(), sos_np)[0]
        C, K = x32.shape[0], sos_t.shape[0]
        z = torch.zeros(K, C, 2, dtype=torch.float64)
        return ext.sos_forward(x32.to(torch.float64), sos_t, sos_t, z, z.clone())[0].to(torch.float32)

    g = torch.Generator().manual_seed(1)
    if kind == "cfg4":
        x = 0.1 * torch.randn(2048, 5 * FS, generator=g)
        fs_ = [fx.filter.LoButterworth(5000, order=4, fs=FS), fx.filter.ParametricEQ(1000, q=2.0, gain=3.0, fs=FS),
               fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=FS)]
        for f in fs_:
            f.compute_coefficients()
        sos = np.vstack([f._sos.numpy() for f in fs_])
        sos_call(x[:, :4800], sos)
        t0 = time.perf_counter()
        sos_call(x, sos)
        dt = time.perf_counter() - t0
        return {"Gsamples_s": round(x.numel() / dt / 1e9, 3), "cores": cores, "kind": "reference" if ext else "port",
                "sample": "2048 ch x 5 s, fused K=4 cascade (one sos_forward call)"}
    if kind == "cfg5":
        x = 0.1 * torch.randn(32, 10 * FS, generator=g)
        bank = fx.filter.LogFilterBank(n_bands=32, f_min=20.0, f_max=20000.0, q=1.414, fs=FS)
        bank.compute_coefficients()
        t0 = time.perf_counter()
        torch.stack([sos_call(x, f._sos.numpy()) for f in bank.filters])  # the reference's Python loop + stack
        dt = time.perf_counter() - t0
        return {"G_lane_samples_s": round(32 * x.numel() / dt / 1e9, 3), "cores": cores, "kind": "reference" if ext else "port",
                "sample": "32 bands x 32 ch x 10 s (32 sos_forward calls + stack)"}
    if kind == "cfg3":
        rng = np.random.default_rng(7)
        K = 65536
        ir = rng.standard_normal(K) * np.exp(-np.arange(K) / 8000.0)
        ir = (ir / np.sqrt((ir ** 2).sum())).astype(np.float32)
        x = (0.1 * rng.standard_normal((1, 16, 30 * FS))).astype(np.float32)
        t0 = time.perf_counter()
        oracle.fft_conv1d(x, ir[::-1].copy(), padding=(K - 1, 0))
        dt = time.perf_counter() - t0
        return {"Gsamples_s": round(x.size / dt / 1e9, 3), "cores": 1, "kind": "port",
                "sample": "16 ch x 30 s, numpy restatement of the reference's overlap-save (block = 5K), single thread"}
    return None


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
    out["cpu_baseline"] = cpu_reference("cfg4")
    return {"config": f"cfg4: fused LoButterworth|ParametricEQ|HiShelving, {C} ch x {SECONDS:g} s", **out}


def cfg5():
    out = {}
    C, T, N = 256, int(SECONDS * FS), 32
    x = torch.empty((C, T), device=DEV).normal_(0, 0.1)
    bank = fx.filter.LogFilterBank(n_bands=N, f_min=20.0, f_max=20000.0, q=1.414, fs=FS)
    lanes = N * C * T
    bytes_alg = 4 * lanes * (1 + 1 / N)
    for prec in ("f32", "f64", "auto"):
        _ops.set_default_precision(prec)
        ms, nl = timed(lambda: (bank.reset_state(), bank(x))[1], reps=3)
        out["stack" if prec == "auto" else f"stack_{prec}"] = {
            "ms": round(ms, 3), "G_lane_samples_s": round(lanes / ms / 1e6, 1), "GBps": round(bytes_alg / ms / 1e6, 1),
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
    for prec in ("f32", "f64", "auto"):
        _ops.set_default_precision(prec)
        ms, nl = timed(run_sum, reps=3)
        out["sum" if prec == "auto" else f"sum_{prec}"] = {
            "ms": round(ms, 3), "Gsamples_s": round(C * T / ms / 1e6, 1), "G_lane_samples_s": round(8 * C * T / ms / 1e6, 1),
            "GBps": round(8 * C * T / ms / 1e6, 1), "frac_hbm": round(8 * C * T / ms / 1e6 / PEAK, 3), "launches": nl}
    # BASELINE configs[4] read literally: 32 biquads added over 256 channels (8192 biquad lanes)
    x32 = x[:256]
    fl32 = [fx.filter.BiquadBPF(20.0 * (1000.0 ** (i / 31.0)), 1.414, FS) for i in range(32)]
    comb32 = fx.filter._base.ParallelFilterCombination(*fl32)
    def run_sum32():
        for f in fl32:
            f.reset_state()
        return comb32(x32)
    for prec in ("f32", "auto"):
        _ops.set_default_precision(prec)
        ms, nl = timed(run_sum32, reps=3)
        out["sum32" if prec == "auto" else f"sum32_{prec}"] = {
            "ms": round(ms, 3), "Gsamples_s": round(256 * T / ms / 1e6, 1), "G_lane_samples_s": round(32 * 256 * T / ms / 1e6, 1),
            "GBps": round(8 * 256 * T / ms / 1e6, 1), "launches": nl}
    for f in fl:
        f.reset_state()
    y = comb(x[:2, : 1 << 17])
    want = oracle.filterbank_sum(x[:2, : 1 << 17].cpu().numpy(), np.stack([f._sos.numpy() for f in fl]))
    out["sum"]["parity_rel_err"] = rel(y.cpu().numpy(), want)
    out["cpu_baseline"] = cpu_reference("cfg5")
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
    out["cpu_baseline"] = cpu_reference("cfg3")
    return {"config": f"cfg3: FIR overlap-save, 65536-tap IR, {C} ch x {SECONDS:g} s", **out}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg4", "cfg5", "cfg3"]
    for name in which:
        print(json.dumps(globals()[name]()), flush=True)
        torch.cuda.empty_cache()
