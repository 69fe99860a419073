"""Wave pipe / chain fuser / parallel combination on CPU tensors -- the reference's
tests/test_chain_fusion.py, test_fused.py, test_filter_base.py, test_filterbank.py."""
from __future__ import annotations

import numpy as np
import pytest
import scipy.signal as sps
import torch
import torch.nn as nn

import torchfx_b200 as fx
from conftest import golden
from torchfx_b200.filter import FusedSOSCascade, HiButterworth, LoButterworth
from torchfx_b200.filter._base import ParallelFilterCombination

FS = 44100


def _chain():
    return [LoButterworth(4000, order=4), fx.filter.ParametricEQ(1000, 2.0, 3.0), HiButterworth(200, order=2)]


def _sequential(x, filters):
    y = x.double().numpy()
    for f in filters:
        y = sps.sosfilt(f._sos.numpy(), y, axis=-1)
    return y


def test_cfg1_golden_through_wave_pipe():
    g = golden("cfg1_lobutter4_mono.npz")
    y = (fx.Wave(torch.from_numpy(g["x"]), 48000) | LoButterworth(cutoff=5000, order=4)).ys
    assert y.dtype == torch.float32 and y.shape == (1, 48000)
    assert np.abs(y.numpy() - g["y"]).max() <= 2e-7 * np.abs(g["y"]).max()


def test_cfg4_golden_deferred_chain_is_fused():
    g = golden("cfg4_chain.npz")
    w = fx.Wave(torch.from_numpy(g["x"]), 48000)
    w = w | LoButterworth(5000, order=4) | fx.filter.ParametricEQ(1000, q=2.0, gain=3.0) | fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db")
    assert len(w._pipeline) == 3  # nothing computed yet
    plan = w._plan()
    assert len(plan) == 1 and isinstance(plan[0], FusedSOSCascade) and plan[0]._num_sections == 4
    np.testing.assert_allclose(plan[0]._sos.numpy(), g["sos"], rtol=1e-13)
    y = w.ys
    assert w._pipeline == []
    assert np.abs(y.numpy() - g["y"]).max() <= 3e-7 * np.abs(g["y"]).max()


def test_three_spellings_of_a_chain_agree():
    torch.manual_seed(42)
    x = torch.randn(2, FS)
    ref = _sequential(x, [f for f in _chain() if (setattr(f, "fs", FS), f.compute_coefficients())])
    a, b, c = _chain(), _chain(), _chain()
    y1 = (fx.Wave(x, FS) | a[0] | a[1] | a[2]).ys
    y2 = (fx.Wave(x, FS) | (b[0] | b[1] | b[2])).ys
    y3 = (fx.Wave(x, FS) | nn.Sequential(*c)).ys
    for y in (y1, y2, y3):
        np.testing.assert_allclose(y.numpy(), ref, atol=2e-6)
    assert isinstance(b[0] | b[1], fx.FilterChain) and len((b[0] | b[1]) | b[2]) == 3


def test_gain_splits_runs_and_laziness():
    torch.manual_seed(3)
    x = torch.randn(2, 8000)
    f = [LoButterworth(3000, order=2), HiButterworth(100, order=2), fx.Gain(0.5), LoButterworth(2000, order=2), HiButterworth(50, order=2)]
    w = fx.Wave(x, FS)
    for m in f:
        w = w | m
    import torchfx_b200.wave as wave_mod

    # reference structure (wave.py:227-233): the gain breaks the IIR run
    wave_mod.FOLD_GAIN = False
    try:
        kinds = [type(m).__name__ for m in w._plan()]
        assert kinds == ["FusedSOSCascade", "Gain", "FusedSOSCascade"]
        y_unfolded = (fx.Wave(x, FS) | f[0] | f[1] | f[2] | f[3] | f[4]).ys
    finally:
        wave_mod.FOLD_GAIN = True
    # default: the gain is folded into the coefficients, one cascade of all four filters
    plan = w._plan()
    assert [type(m).__name__ for m in plan] == ["FusedSOSCascade"]
    assert plan[0]._num_sections == 4 and plan[0].gain == 0.5
    assert f[0]._state_x is None  # lazy: nothing ran
    y = w.ys
    np.testing.assert_allclose(y.numpy(), y_unfolded.numpy(), atol=2e-6)
    ref = _sequential(x, f[:2]) * 0.5
    ref = sps.sosfilt(f[4]._sos.numpy(), sps.sosfilt(f[3]._sos.numpy(), ref, axis=-1), axis=-1)
    np.testing.assert_allclose(y.numpy(), ref, atol=2e-6)
    with pytest.raises(TypeError, match="Expected nn.Module"):
        fx.Wave(x, FS) | 3


def test_gain_folding_rules():
    torch.manual_seed(4)
    x = torch.randn(2, 4000)
    lo, hi = LoButterworth(3000, order=2), HiButterworth(100, order=2)
    lo2, hi2 = LoButterworth(2500, order=2), HiButterworth(60, order=2)
    # leading / trailing / dB gains fold into runs of >= 2 filters and merge the runs around them
    w = fx.Wave(x, FS) | fx.Gain(2.0) | lo | hi | fx.Gain(-6.0, gain_type="db") | lo2 | hi2 | fx.Gain(0.25)
    plan = w._plan()
    assert [type(m).__name__ for m in plan] == ["FusedSOSCascade"]
    g = 2.0 * 10 ** (-6.0 / 20) * 0.25
    assert abs(plan[0].gain - g) < 1e-12 and plan[0]._num_sections == 4
    ref = _sequential(x, [lo, hi, lo2, hi2]) * g
    np.testing.assert_allclose(w.ys.numpy(), ref, atol=2e-6)
    # a clamping gain never folds
    w = fx.Wave(x, FS) | lo | fx.Gain(4.0, clamp=True) | hi
    assert [type(m).__name__ for m in w._plan()] == ["LoButterworth", "Gain", "HiButterworth"]
    w = fx.Wave(x, FS) | fx.Gain(0.5) | LoButterworth(3000, order=2)
    assert [type(m).__name__ for m in w._plan()] == ["Gain", "LoButterworth"]
    # a lone filter is a stateful step in the reference (wave.py:227-233): it is never absorbed, and it separates
    w = fx.Wave(x, FS) | lo | hi | fx.Gain(0.5) | lo2
    assert [type(m).__name__ for m in w._plan()] == ["FusedSOSCascade", "LoButterworth"]
    w = fx.Wave(x, FS) | lo | fx.Gain(0.5) | hi
    assert [type(m).__name__ for m in w._plan()] == ["LoButterworth", "Gain", "HiButterworth"]


def test_gain_between_lone_filters_keeps_their_state():
    """``Wave(chunk) | lp | Gain | hp`` per chunk must continue the stream as in the reference --
    lp and hp run as the modules themselves and carry their DF1 state between materialisations."""
    torch.manual_seed(5)
    x = torch.randn(2, 6000)
    lp, hp = LoButterworth(3000, order=4), HiButterworth(100, order=2)
    g = fx.Gain(0.5)
    pieces = [(fx.Wave(x[:, a:b], FS) | lp | g | hp).ys for a, b in ((0, 2500), (2500, 6000))]
    assert lp._state_x is not None and hp._state_x is not None
    ref = sps.sosfilt(hp._sos.numpy(), 0.5 * sps.sosfilt(lp._sos.numpy(), x.numpy().astype(np.float64), axis=-1), axis=-1)
    np.testing.assert_allclose(torch.cat(pieces, dim=1).numpy(), ref, atol=2e-6)


def test_fused_construction_errors_and_from_chain():
    with pytest.raises(ValueError, match="at least one"):
        FusedSOSCascade()
    with pytest.raises(TypeError, match="Expected filter with SOS coefficients"):
        FusedSOSCascade(object())
    with pytest.raises(ValueError, match="different sample rates"):
        FusedSOSCascade(LoButterworth(2000, order=4, fs=44100), LoButterworth(2000, order=4, fs=48000))
    with pytest.raises(ValueError, match="no sampling frequency"):
        FusedSOSCascade(LoButterworth(2000, order=4))
    f = LoButterworth(2000, order=4, fs=FS)
    assert f._sos is None
    fused = FusedSOSCascade(f)
    assert f._sos is not None and fused._num_sections == 2 and fused.fs == FS and fused._sos.dtype == torch.float64
    chain = nn.Sequential(LoButterworth(2000, order=4, fs=FS), HiButterworth(200, order=2, fs=FS))
    assert FusedSOSCascade.from_chain(chain)._num_sections == 3
    assert FusedSOSCascade.from_chain(f)._num_sections == 2
    with pytest.raises(TypeError, match="Expected nn.Sequential or IIR/Biquad"):
        FusedSOSCascade.from_chain(nn.ReLU())
    with pytest.raises(ValueError, match="No IIR/Biquad filters"):
        FusedSOSCascade.from_chain(nn.Sequential(nn.Identity()))


@pytest.mark.parametrize("order", [6, 12, 20])
def test_fused_high_orders_and_chunking(order):
    torch.manual_seed(order)
    x = torch.randn(2, 20000, dtype=torch.float64)
    f = LoButterworth(3000, order=order, fs=FS)
    fused = FusedSOSCascade(f)
    y = fused(x)
    np.testing.assert_allclose(y.numpy(), sps.sosfilt(f._sos.numpy(), x.numpy(), axis=-1), atol=1e-9)
    fused.reset_state()
    parts = torch.cat([fused(x[:, :7000]), fused(x[:, 7000:])], dim=1)
    torch.testing.assert_close(parts, y, atol=1e-12, rtol=0)
    assert fused._stateful and y.dtype == torch.float64


def test_parallel_combination_semantics():
    g = golden("parallel.npz")
    x = torch.from_numpy(g["x"])
    f1, f2, f3 = fx.filter.BiquadBPF(500, 1.414, 48000), fx.filter.BiquadBPF(2000, 1.414, 48000), LoButterworth(300, order=2, fs=48000)
    comb = f1 + f2 + f3
    assert isinstance(comb, ParallelFilterCombination) and comb.filters[1] is f3
    y = comb(x)
    assert np.abs(y.numpy() - g["comb"]).max() <= 3e-7 * np.abs(g["comb"]).max()
    # fs propagation: children keep an fs they already have (reference tests/test_filter_base.py:88-130)
    a, b = LoButterworth(1000, order=2), HiButterworth(100, order=2, fs=22050)
    p = ParallelFilterCombination(a, b, fs=48000)
    assert a.fs == 48000 and b.fs == 22050 and p.fs == 48000
    assert not p._has_computed_coeff
    p.compute_coefficients()
    assert p._has_computed_coeff
    with pytest.raises(AssertionError):
        a + 3
    # nested (f1 + f2) | f3 through a Wave
    torch.manual_seed(0)
    xs = torch.randn(2, 5000)
    q1, q2, q3 = LoButterworth(1000, order=2), HiButterworth(200, order=2), LoButterworth(4000, order=2)
    y = (fx.Wave(xs, FS) | ((q1 + q2) | q3)).ys
    inner = sps.sosfilt(q1._sos.numpy(), xs.double().numpy(), axis=-1) + sps.sosfilt(q2._sos.numpy(), xs.double().numpy(), axis=-1)
    np.testing.assert_allclose(y.numpy(), sps.sosfilt(q3._sos.numpy(), inner, axis=-1), atol=3e-6)


def test_logfilterbank_cpu():
    g = golden("parallel.npz")
    bank = fx.filter.LogFilterBank(n_bands=8, f_min=100.0, f_max=8000.0, fs=48000)
    y = bank(torch.from_numpy(g["x"]))
    assert y.shape == (8, 2, 4096)
    assert np.abs(y.numpy() - g["bank"]).max() <= 3e-7 * np.abs(g["bank"]).max()
    cf = bank.center_frequencies
    assert cf[0] == pytest.approx(100.0) and cf[-1] == pytest.approx(8000.0)
    ratios = [cf[i + 1] / cf[i] for i in range(7)]
    assert max(ratios) - min(ratios) < 1e-9  # log spacing
    with pytest.raises(AssertionError):
        fx.filter.LogFilterBank(1)
    with pytest.raises(AssertionError):
        fx.filter.LogFilterBank(4, f_min=100, f_max=50)
    b2 = fx.filter.LogFilterBank(4)
    b2.fs = 16000
    assert all(f.fs == 16000 for f in b2.filters)


def test_wave_helpers():
    x = torch.randn(2, 1000)
    w = fx.Wave(x, 8000)
    assert len(w) == 1000 and w.channels() == 2 and w.duration("sec") == 0.125 and w.duration("ms") == 125.0
    assert w.get_channel(1).ys.shape == (1000,)
    m = fx.Wave.merge([w, fx.Wave(torch.randn(2, 400), 8000)])
    assert m.ys.shape == (2, 1000)
    s = fx.Wave.merge([w, w], split_channels=True)
    assert s.ys.shape == (4, 1000)
    with pytest.raises(ValueError, match="Sampling frequency mismatch"):
        fx.Wave.merge([w, fx.Wave(x, 16000)])
    with pytest.raises(ValueError, match="No waves"):
        fx.Wave.merge([])
    t = w.transform(lambda t: t * 2)
    torch.testing.assert_close(t.ys, x * 2)
