"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/torchfx_ext.so):

    make -C oracle ref && python oracle/make_golden.py

Every vector is produced by the reference's own public API / `_ops` wrappers on seeded
inputs; inputs are stored next to the outputs so the fixtures do not depend on torch's RNG.
The reference ships no golden files of its own -- these play that role.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') != HERE]
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main() -> None:
    fx = ref_loader.import_reference()
    from torchfx import _ops as ref_ops
    from torchfx.filter import _fftconv as ref_fftconv

    os.makedirs(OUT, exist_ok=True)
    F = fx.filter
    g = torch.Generator().manual_seed(20260101)

    def randn(*shape, dtype=torch.float32, scale=1.0):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale).to(dtype)

    # ---- cfg1: BASELINE.json configs[0] --------------------------------------------------
    torch.manual_seed(0)
    x = torch.randn(1, 48000)
    lp = F.LoButterworth(cutoff=5000, order=4)
    y = (fx.Wave(x, 48000) | lp).ys
    np.savez_compressed(os.path.join(OUT, "cfg1_lobutter4_mono.npz"), x=x.numpy(), y=y.numpy(), sos=lp._sos.numpy(), fs=48000)

    # ---- K=4 cascade (cfg2 filter), chunked with state carry ------------------------------
    x = randn(4, 8192, scale=0.1)
    f = F.LoButterworth(cutoff=5000, order=8, fs=48000)
    y_full = f(x)
    sx_full, sy_full = f._state_x.clone(), f._state_y.clone()
    f2 = F.LoButterworth(cutoff=5000, order=8, fs=48000)
    ya = f2(x[:, :3000])
    sx_mid, sy_mid = f2._state_x.clone(), f2._state_y.clone()
    yb = f2(x[:, 3000:])
    assert torch.equal(torch.cat([ya, yb], 1), y_full) or torch.allclose(torch.cat([ya, yb], 1), y_full, atol=1e-7)
    np.savez_compressed(
        os.path.join(OUT, "sos_k4_chunked.npz"), x=x.numpy(), y=y_full.numpy(), sos=f._sos.numpy(),
        state_x=sx_full.numpy(), state_y=sy_full.numpy(), state_x_mid=sx_mid.numpy(), state_y_mid=sy_mid.numpy(), split=3000,
    )

    # ---- _ops wrappers with a non-zero initial state, f64 ----------------------------------
    x = randn(3, 2048, dtype=torch.float64)
    sos = torch.from_numpy(__import__("scipy.signal").signal.cheby1(6, 0.5, 0.3, output="sos"))
    sx0 = randn(3, 3, 2, dtype=torch.float64)
    sy0 = randn(3, 3, 2, dtype=torch.float64)
    y, sx1, sy1 = ref_ops.parallel_iir_forward(x, sos, sx0, sy0)
    b = sos[0, :3].clone()
    a = sos[0, 3:].clone()
    yb, bsx, bsy = ref_ops.biquad_forward(x, b, a, sx0[0], sy0[0])
    np.savez_compressed(
        os.path.join(OUT, "ops_state_f64.npz"), x=x.numpy(), sos=sos.numpy(), sx0=sx0.numpy(), sy0=sy0.numpy(),
        y=y.numpy(), sx1=sx1.numpy(), sy1=sy1.numpy(), yb=yb.numpy(), bsx=bsx.numpy(), bsy=bsy.numpy(),
    )

    # ---- cfg4 chain, deferred + auto-fused ------------------------------------------------
    x = randn(2, 8192, scale=0.1)
    chain = [F.LoButterworth(5000, order=4), F.ParametricEQ(1000, q=2.0, gain=3.0), F.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db")]
    w = fx.Wave(x, 48000)
    for c in chain:
        w = w | c
    y = w.ys
    sos_cat = torch.cat([c._sos for c in chain], 0)
    np.savez_compressed(os.path.join(OUT, "cfg4_chain.npz"), x=x.numpy(), y=y.numpy(), sos=sos_cat.numpy(), fs=48000)

    # ---- coefficient designs ----------------------------------------------------------------
    designs = {}
    fs = 44100
    for name, ctor in {
        "BiquadLPF": lambda: F.BiquadLPF(1000, 0.707, fs), "BiquadHPF": lambda: F.BiquadHPF(1000, 0.707, fs),
        "BiquadNotch": lambda: F.BiquadNotch(1000, 5.0, fs), "BiquadBPF": lambda: F.BiquadBPF(1000, 1.414, fs),
        "BiquadBPFPeak": lambda: F.BiquadBPFPeak(1000, 1.414, fs), "BiquadAllPass": lambda: F.BiquadAllPass(1000, 0.707, fs),
        "HiShelving": lambda: F.HiShelving(8000, 0.707, 2.0, "db", fs), "LoShelving": lambda: F.LoShelving(200, 0.707, 1.5, "linear", fs),
        "ParametricEQ": lambda: F.ParametricEQ(1000, 2.0, 3.0, fs), "Peaking": lambda: F.Peaking(2000, 1.0, 2.0, "linear", fs),
        "Notch": lambda: F.Notch(60, 10.0, fs), "AllPass": lambda: F.AllPass(500, 0.9, fs),
        "LoButterworth": lambda: F.LoButterworth(2000, fs=fs), "HiButterworth": lambda: F.HiButterworth(200, order=3, fs=fs),
        "LoButterworth_db": lambda: F.LoButterworth(2000, order=24, order_scale="db", fs=fs),
        "HiChebyshev1": lambda: F.HiChebyshev1(300, order=4, ripple=0.5, fs=fs), "LoChebyshev2": lambda: F.LoChebyshev2(3000, order=5, ripple=30, fs=fs),
        "LoElliptic": lambda: F.LoElliptic(3000, order=4, fs=fs), "HiLinkwitzRiley": lambda: F.HiLinkwitzRiley(1500, order=4, fs=fs),
    }.items():
        flt = ctor()
        flt.compute_coefficients()
        designs[name] = flt._sos.numpy()
    np.savez_compressed(os.path.join(OUT, "designs.npz"), **designs)

    # ---- parallel semantics -------------------------------------------------------------------
    x = randn(2, 4096, scale=0.1)
    bank = F.LogFilterBank(n_bands=8, f_min=100.0, f_max=8000.0, fs=48000)
    yb = bank(x)
    comb = F.BiquadBPF(500, 1.414, 48000) + F.BiquadBPF(2000, 1.414, 48000) + F.LoButterworth(300, order=2, fs=48000)
    ys = comb(x)
    np.savez_compressed(
        os.path.join(OUT, "parallel.npz"), x=x.numpy(), bank=yb.numpy(), bank_sos=np.stack([f._sos.numpy() for f in bank.filters]),
        comb=ys.numpy(), fs=48000,
    )

    # ---- FIR / overlap-save -------------------------------------------------------------------
    x = randn(2, 4096, scale=0.1)
    taps = randn(101).numpy()
    yf = F.FIR(taps)(x)
    yd = F.FIR(taps, conv_mode="direct")(x)
    des = F.DesignableFIR(cutoff=3000.0, num_taps=63, fs=48000)
    ydes = des(x)
    kern = randn(1, 1, 64)
    xc = randn(1, 2, 1000)
    yc = ref_fftconv.fft_conv1d(xc, kern)
    yc_pad = ref_fftconv.fft_conv1d(xc, kern, padding=(5, 10))
    np.savez_compressed(
        os.path.join(OUT, "fir.npz"), x=x.numpy(), taps=taps, y_fft=yf.numpy(), y_direct=yd.numpy(), des_b=np.asarray(des.b),
        y_des=ydes.numpy(), kern=kern.numpy(), xc=xc.numpy(), yc=yc.numpy(), yc_pad=yc_pad.numpy(),
    )

    # ---- delay line ---------------------------------------------------------------------------
    x = randn(2, 2048)
    y32 = ref_ops.delay_line_forward(x, 100, 0.5, 0.8)
    x64 = randn(2, 2048, dtype=torch.float64)
    y64 = ref_ops.delay_line_forward(x64, 333, 0.7, 0.25)
    np.savez_compressed(os.path.join(OUT, "delay.npz"), x=x.numpy(), y=y32.numpy(), x64=x64.numpy(), y64=y64.numpy())

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
