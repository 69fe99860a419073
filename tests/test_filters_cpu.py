"""Host-side surface on CPU tensors: filter classes, coefficient designs, state contract,
shape/dtype handling -- the behaviours the reference's tests pin (tests/test_iir.py,
test_iir_gaps.py, test_biquad.py, test_ops_dispatch.py), checked against scipy, the closed
forms and the reference-generated golden designs."""
from __future__ import annotations

import math

import numpy as np
import pytest
import scipy.signal as sps
import torch

import torchfx_b200 as fx
from conftest import golden
from torchfx_b200 import _ops

F = fx.filter
FS = 44100


def test_designs_match_reference_golden():
    g = golden("designs.npz")
    ctors = {
        "BiquadLPF": lambda: F.BiquadLPF(1000, 0.707, FS), "BiquadHPF": lambda: F.BiquadHPF(1000, 0.707, FS),
        "BiquadNotch": lambda: F.BiquadNotch(1000, 5.0, FS), "BiquadBPF": lambda: F.BiquadBPF(1000, 1.414, FS),
        "BiquadBPFPeak": lambda: F.BiquadBPFPeak(1000, 1.414, FS), "BiquadAllPass": lambda: F.BiquadAllPass(1000, 0.707, FS),
        "HiShelving": lambda: F.HiShelving(8000, 0.707, 2.0, "db", FS), "LoShelving": lambda: F.LoShelving(200, 0.707, 1.5, "linear", FS),
        "ParametricEQ": lambda: F.ParametricEQ(1000, 2.0, 3.0, FS), "Peaking": lambda: F.Peaking(2000, 1.0, 2.0, "linear", FS),
        "Notch": lambda: F.Notch(60, 10.0, FS), "AllPass": lambda: F.AllPass(500, 0.9, FS),
        "LoButterworth": lambda: F.LoButterworth(2000, fs=FS), "HiButterworth": lambda: F.HiButterworth(200, order=3, fs=FS),
        "LoButterworth_db": lambda: F.LoButterworth(2000, order=24, order_scale="db", fs=FS),
        "HiChebyshev1": lambda: F.HiChebyshev1(300, order=4, ripple=0.5, fs=FS), "LoChebyshev2": lambda: F.LoChebyshev2(3000, order=5, ripple=30, fs=FS),
        "LoElliptic": lambda: F.LoElliptic(3000, order=4, fs=FS), "HiLinkwitzRiley": lambda: F.HiLinkwitzRiley(1500, order=4, fs=FS),
    }
    assert set(ctors) == set(g.files)
    for name, ctor in ctors.items():
        f = ctor()
        f.compute_coefficients()
        assert f._sos.dtype == torch.float64 and f._sos.shape[1] == 6
        np.testing.assert_allclose(f._sos.numpy(), g[name], rtol=1e-12, atol=1e-15, err_msg=name)


def test_defaults_and_order_scale():
    assert F.LoButterworth(1000).order == 5 and F.HiButterworth(1000).order == 5  # reference iir.py:868,918
    assert F.LoButterworth(1000, order=24, order_scale="db").order == 4
    f = F.Butterworth("lowpass", 1000, fs=FS)
    f.compute_coefficients()
    np.testing.assert_allclose(f._sos.numpy(), sps.butter(4, 1000 / (FS / 2), output="sos"))
    with pytest.raises(ValueError, match="positive even"):
        F.LinkwitzRiley("lowpass", 1000, order=3)
    with pytest.raises(ValueError, match="positive even"):
        F.LinkwitzRiley("lowpass", 1000, order=0)
    lr = F.LoLinkwitzRiley(1000, order=4, fs=FS)
    lr.compute_coefficients()
    half = sps.butter(2, 1000 / (FS / 2), output="sos")
    np.testing.assert_allclose(lr._sos.numpy(), np.vstack([half, half]))


def test_biquad_lpf_closed_form():
    f = F.BiquadLPF(1000, 0.707, FS)
    f.compute_coefficients()
    w0 = 2 * math.pi * 1000 / FS
    alpha = math.sin(w0) / (2 * 0.707)
    a0 = 1 + alpha
    want = [(1 - math.cos(w0)) / 2 / a0, (1 - math.cos(w0)) / a0, (1 - math.cos(w0)) / 2 / a0, 1.0, -2 * math.cos(w0) / a0, (1 - alpha) / a0]
    np.testing.assert_allclose(f._sos.numpy()[0], want, rtol=1e-10)
    assert f.a[0] == 1.0 and f.b.shape == (3,)


def test_no_fs_raises_like_reference():
    with pytest.raises(ValueError, match="[Ss]ample rate"):
        F.LoButterworth(1000)(torch.randn(2, 100))
    with pytest.raises(ValueError, match="[Ss]ample rate"):
        F.BiquadLPF(1000, 0.7)(torch.randn(2, 100))
    with pytest.raises(ValueError, match="[Ss]ample rate"):
        F.LogFilterBank(4)(torch.randn(2, 100))


@pytest.mark.parametrize("shape", [(4410,), (2, 4410), (3, 2, 4410)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_shapes_dtypes_and_scipy(shape, dtype):
    torch.manual_seed(0)
    x = torch.randn(*shape, dtype=dtype)
    f = F.LoChebyshev1(2000, order=6, ripple=0.5, fs=FS)
    y = f(x)
    assert y.shape == x.shape and y.dtype == dtype
    ref = sps.sosfilt(f._sos.numpy(), x.double().numpy(), axis=-1)
    np.testing.assert_allclose(y.double().numpy(), ref, atol=1e-5 if dtype == torch.float32 else 1e-11)
    C = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
    assert f._state_x.shape == (3, C, 2) and f._state_x.dtype == torch.float64


def test_state_carry_reset_and_channel_change():
    torch.manual_seed(1)
    x = torch.randn(2, 6000, dtype=torch.float64)
    f = F.HiButterworth(300, order=4, fs=FS)
    whole = F.HiButterworth(300, order=4, fs=FS)(x)
    parts = torch.cat([f(x[:, :1000]), f(x[:, 1000:1001]), f(x[:, 1001:])], dim=1)
    torch.testing.assert_close(parts, whole, atol=1e-12, rtol=0)
    f.reset_state()
    assert f._state_x is None and f._sos is None  # IIR.reset_state drops the design too (reference iir.py:255-265)
    torch.testing.assert_close(f(x), whole, atol=1e-12, rtol=0)
    y4 = f(torch.randn(4, 100, dtype=torch.float64))  # channel count change re-allocates zero state
    assert f._state_x.shape == (2, 4, 2) and y4.shape == (4, 100)
    b = F.BiquadNotch(1000, 5.0, FS)
    b(x)
    b.reset_state()
    assert b._state_x is None and b._sos is not None  # Biquad.reset_state keeps it (biquad.py:198-206)


def test_ops_wrappers_match_reference_contract():
    g = golden("ops_state_f64.npz")
    x = torch.from_numpy(g["x"])
    sos = torch.from_numpy(g["sos"])
    sx0, sy0 = torch.from_numpy(g["sx0"]), torch.from_numpy(g["sy0"])
    y, sx1, sy1 = _ops.parallel_iir_forward(x, sos, sx0, sy0)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(sx1.numpy(), g["sx1"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(sy1.numpy(), g["sy1"], rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(sx0.numpy(), g["sx0"])  # functional: inputs untouched
    yb, bsx, bsy = _ops.biquad_forward(x, sos[0, :3], sos[0, 3:], sx0[0], sy0[0])
    np.testing.assert_allclose(yb.numpy(), g["yb"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(bsx.numpy(), g["bsx"], rtol=1e-10, atol=1e-12)
    # passthrough + default state shapes (reference tests/test_ops_dispatch.py:48-60,102-115)
    ident = torch.tensor([[1.0, 0, 0, 1, 0, 0]] * 2, dtype=torch.float64)
    y, sx, sy = _ops.parallel_iir_forward(x, ident, None, None)
    torch.testing.assert_close(y, x)
    assert sx.shape == (2, 3, 2) and sy.shape == (2, 3, 2)
    y1, s1, s2 = _ops.biquad_forward(x[0], ident[0, :3], ident[0, 3:], None, None)
    assert y1.shape == x[0].shape and s1.shape == (1, 2)
    # the torchfx_ext-shaped object
    y3, _, _ = fx.torchfx_ext.sos_forward(x, sos, sos, sx0, sy0)
    np.testing.assert_allclose(y3.numpy(), g["y"], rtol=1e-10, atol=1e-12)


def test_delay_line_and_reverb_cpu():
    g = golden("delay.npz")
    y = _ops.delay_line_forward(torch.from_numpy(g["x"]), 100, 0.5, 0.8)
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-6, atol=1e-7)
    y64 = _ops.delay_line_forward(torch.from_numpy(g["x64"]), 333, 0.7, 0.25)
    np.testing.assert_allclose(y64.numpy(), g["y64"], rtol=1e-14)
    x = torch.arange(10, dtype=torch.float64).unsqueeze(0)
    y = _ops.delay_line_forward(x, 3, 0.5, 1.0)
    want = x.clone()
    want[0, 3:] = x[0, 3:] + 0.5 * x[0, :-3]
    torch.testing.assert_close(y, want)
    short = torch.randn(2, 50)
    assert _ops.delay_line_forward(short, 100, 0.5, 0.5) is short
    assert _ops.delay_line_forward(torch.randn(512), 50, 0.5, 0.5).shape == (512,)
    rv = fx.Reverb(delay=100, decay=0.5, mix=0.8)
    np.testing.assert_allclose(rv(torch.from_numpy(g["x"])).numpy(), g["y"], rtol=1e-6, atol=1e-7)


def test_fir_cpu_matches_golden():
    g = golden("fir.npz")
    x = torch.from_numpy(g["x"])
    for mode, key in (("fft", "y_fft"), ("direct", "y_direct"), ("auto", "y_fft")):
        y = F.FIR(g["taps"], conv_mode=mode)(x)
        np.testing.assert_allclose(y.numpy(), g[key], atol=2e-6 * np.abs(g[key]).max())
    from torchfx_b200.filter._fftconv import fft_conv1d

    yc = fft_conv1d(torch.from_numpy(g["xc"]), torch.from_numpy(g["kern"]), padding=(5, 10))
    np.testing.assert_allclose(yc.numpy(), g["yc_pad"], atol=2e-6 * np.abs(g["yc_pad"]).max())
    assert F.FIR(g["taps"]).a == [1.0]
    sd = F.FIR(g["taps"]).state_dict()
    assert list(sd) == ["kernel"] and sd["kernel"].shape == (1, 1, 101)


def test_exports_match_reference_filter_namespace():
    names = ["AllPass", "Biquad", "BiquadAllPass", "BiquadBPF", "BiquadBPFPeak", "BiquadHPF", "BiquadLPF", "BiquadNotch",
             "Butterworth", "Chebyshev1", "Chebyshev2", "DesignableFIR", "Elliptic", "FIR", "FusedSOSCascade", "HiButterworth",
             "HiChebyshev1", "HiChebyshev2", "HiElliptic", "HiLinkwitzRiley", "HiShelving", "IIR", "LinkwitzRiley", "LoButterworth",
             "LoChebyshev1", "LoChebyshev2", "LoElliptic", "LoLinkwitzRiley", "LogFilterBank", "LoShelving", "Notch", "ParametricEQ",
             "Peaking"]
    assert sorted(F.__all__) == sorted(names)
    for n in names:
        assert hasattr(F, n)
