"""GPU parity of the channel-tile cascade kernel (sos_tile.cu: lanes = 32 consecutive channels, cp.async tiles) -- the
path every BASELINE config takes (>= 26 channels) -- against the oracle, and A/B against the stream-per-lane kernel
(TFX_NO_TILE) on the same inputs; float64 I/O; the mixed-precision instantiations."""
from __future__ import annotations

import numpy as np
import pytest
import scipy.signal as sps
import torch

from conftest import rel_to_max
from oracle import oracle
from torchfx_b200 import _native, _ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_F32 = 1e-5
TOL_F64REC = 1e-6


def run(x_np, sos_np, sx=None, sy=None, **kw):
    x = torch.from_numpy(np.ascontiguousarray(x_np)).to(DEV)
    K, C = sos_np.shape[0], x.shape[0]
    stx = torch.zeros(K, C, 2, dtype=torch.float64, device=DEV) if sx is None else torch.from_numpy(sx).to(DEV)
    sty = torch.zeros(K, C, 2, dtype=torch.float64, device=DEV) if sy is None else torch.from_numpy(sy).to(DEV)
    y = _ops.sos_cascade_(x, torch.from_numpy(sos_np), stx, sty, **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy(), stx.cpu().numpy(), sty.cpu().numpy(), x


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6, 7, 8, 12])
def test_every_section_count(K):
    rng = np.random.default_rng(100 + K)
    x = (0.1 * rng.standard_normal((64, 12000))).astype(np.float32)
    sos = sps.butter(2 * K, 0.21, output="sos")
    want, wsx, wsy = oracle.sos_cascade(x, sos)
    for precision, tol in (("f32", TOL_F32), ("f64", TOL_F64REC)):
        y, sx, sy, xt = run(x, sos, precision=precision)
        assert rel_to_max(y, want) < tol, (K, precision)
        np.testing.assert_allclose(sx, wsx, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(sy, wsy, rtol=1e-3, atol=tol * 10 * max(np.abs(wsy).max(), 1e-30))


@pytest.mark.parametrize("shape", [(32, 4), (32, 8), (64, 60), (64, 64), (64, 68), (96, 128), (96, 132), (128, 1000), (52, 260), (27, 2052)])
def test_short_and_ragged_with_state(shape):
    """T around the 64-sample chunk and the 2-chunk tracked tail; channel counts that leave
    dead lanes; non-zero initial DF1 state in and out."""
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    sos = sps.cheby1(4, 1.0, 0.3, output="sos")
    sx0 = rng.standard_normal((2, shape[0], 2))
    sy0 = rng.standard_normal((2, shape[0], 2))
    want, wsx, wsy = oracle.sos_cascade(x, sos, sx0, sy0)
    y, sx, sy, xt = run(x, sos, sx0.copy(), sy0.copy(), precision="f64")
    assert rel_to_max(y, want) < TOL_F64REC
    np.testing.assert_allclose(sx, wsx, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(sy, wsy, rtol=1e-5, atol=1e-5)


def test_time_split_many_channels_matches_generic_and_oracle():
    rng = np.random.default_rng(7)
    x = (0.1 * rng.standard_normal((64, 1 << 19))).astype(np.float32)
    sos = sps.butter(8, 5000 / 24000, output="sos")
    pick = [0, 31, 32, 63]
    want, _, wsy = oracle.sos_cascade(x[pick], sos)
    y_t, _, sy_t, _ = run(x, sos, precision="f32")
    y_g, _, sy_g, _ = run(x, sos, precision="f32", no_tile=True)  # stream-per-lane kernel; y_t: channel-tile kernel
    y_n, _, _, _ = run(x, sos, precision="f32", no_split=True)
    assert rel_to_max(y_t[pick], want) < TOL_F32
    assert rel_to_max(y_g[pick], want) < TOL_F32
    assert rel_to_max(y_t, y_n) < 2e-6   # split vs unsplit: warm-up truncation only
    assert rel_to_max(y_t, y_g) < 2e-6   # the two kernels agree
    np.testing.assert_allclose(sy_t, sy_g, rtol=1e-4, atol=1e-6 * np.abs(wsy).max())


def test_chunked_stream_and_in_place():
    rng = np.random.default_rng(9)
    x = (0.1 * rng.standard_normal((64, 100000))).astype(np.float32)
    sos_np = sps.butter(8, 5000 / 24000, output="sos")
    sos = torch.from_numpy(sos_np)
    want, _, wsy = oracle.sos_cascade(x, sos_np)
    xt = torch.from_numpy(x).to(DEV)
    sx = torch.zeros(4, 64, 2, dtype=torch.float64, device=DEV)
    sy = torch.zeros_like(sx)
    for lo, hi in ((0, 30000), (30000, 30004), (30004, 100000)):
        blk = xt[:, lo:hi]
        _ops.sos_cascade_(blk, sos, sx, sy, out=blk)  # in place on a strided view
    assert rel_to_max(xt.cpu().numpy(), want) < TOL_F32
    np.testing.assert_allclose(sy.cpu().numpy(), wsy, rtol=1e-3, atol=1e-5 * np.abs(wsy).max())


def test_f64_io():
    rng = np.random.default_rng(21)
    x = rng.standard_normal((40, 5000))
    x = np.concatenate([x, x[:12]], 0)  # 52 channels
    sos = sps.ellip(6, 0.5, 50, 0.25, output="sos")
    sx0 = rng.standard_normal((3, 52, 2))
    sy0 = rng.standard_normal((3, 52, 2))
    want, wsx, wsy = oracle.sos_cascade(x, sos, sx0, sy0)
    y, sx, sy, xt = run(x, sos, sx0.copy(), sy0.copy())
    np.testing.assert_allclose(y, want, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(sx, wsx, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(sy, wsy, rtol=1e-9, atol=1e-10)


def test_out_of_place_keeps_input():
    rng = np.random.default_rng(33)
    x = (0.1 * rng.standard_normal((32, 70000))).astype(np.float32)
    sos = sps.butter(4, 0.2, output="sos")
    want, _, _ = oracle.sos_cascade(x, sos)
    xt = torch.from_numpy(x).to(DEV)
    y = _ops.sos_cascade_(xt, torch.from_numpy(sos), None, None)
    assert torch.equal(xt.cpu(), torch.from_numpy(x))
    assert rel_to_max(y.cpu().numpy(), want) < TOL_F32


# ---- more shapes: odd T, unaligned rows, long cascades, strided views -------------------------------
@pytest.mark.parametrize("shape", [(32, 1), (32, 2), (32, 3), (64, 63), (64, 64), (64, 65), (96, 129), (128, 1000), (52, 260),
                                   (27, 2052), (33, 4099), (100, 777)])
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_tile_kernel_short_ragged_with_state(shape, precision):
    """Default kernel for >= 26 channels: T around one and two chunks (tracked tail), odd T
    (unaligned rows -> element-wise copies), half-empty channel groups, DF1 state in and out."""
    rng = np.random.default_rng(sum(shape) + 7)
    x = rng.standard_normal(shape).astype(np.float32)
    sos = sps.cheby1(4, 1.0, 0.3, output="sos")
    sx0 = rng.standard_normal((2, shape[0], 2))
    sy0 = rng.standard_normal((2, shape[0], 2))
    want, wsx, wsy = oracle.sos_cascade(x, sos, sx0, sy0)
    y, sx, sy, _ = run(x, sos, sx0.copy(), sy0.copy(), precision=precision)
    tol = TOL_F32 if precision == "f32" else TOL_F64REC
    assert rel_to_max(y, want) < tol
    np.testing.assert_allclose(sx, wsx, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(sy, wsy, rtol=1e-4, atol=tol * 10 * max(np.abs(wsy).max(), 1.0))


@pytest.mark.parametrize("K", [1, 3, 4, 8, 11, 16])
def test_tile_kernel_section_counts_and_f64_io(K):
    rng = np.random.default_rng(900 + K)
    x = rng.standard_normal((40, 30000))
    sos = sps.butter(2 * K, 0.23, output="sos")
    want, wsx, wsy = oracle.sos_cascade(x, sos)
    y, sx, sy, _ = run(x, sos)  # float64 I/O through the tile kernel
    np.testing.assert_allclose(y, want, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(sy, wsy, rtol=1e-8, atol=1e-10)
    xf = x.astype(np.float32)
    wf, _, _ = oracle.sos_cascade(xf, sos)
    yf, _, _, _ = run(xf, sos, precision="f32")
    assert rel_to_max(yf, wf) < TOL_F32


def test_tile_kernel_chunked_in_place_strided_view():
    rng = np.random.default_rng(77)
    big = torch.from_numpy((0.1 * rng.standard_normal((70, 120000))).astype(np.float32)).to(DEV)
    view = big[3:67, 64:100064]  # 64 channels, row stride 120000, aligned start
    x_np = view.cpu().numpy().copy()
    sos_np = sps.butter(8, 5000 / 24000, output="sos")
    sos = torch.from_numpy(sos_np)
    want, _, wsy = oracle.sos_cascade(x_np, sos_np)
    sx = torch.zeros(4, 64, 2, dtype=torch.float64, device=DEV)
    sy = torch.zeros_like(sx)
    for lo, hi in ((0, 40000), (40000, 40004), (40004, 100000)):
        blk = view[:, lo:hi]
        _ops.sos_cascade_(blk, sos, sx, sy, out=blk)
    assert rel_to_max(view.cpu().numpy(), want) < TOL_F32
    np.testing.assert_allclose(sy.cpu().numpy(), wsy, rtol=1e-3, atol=1e-5 * np.abs(wsy).max())
    assert torch.equal(big[0], torch.from_numpy((0.1 * np.random.default_rng(77).standard_normal((70, 120000))).astype(np.float32))[0].to(DEV))


def test_mixed_precision_chain_cfg4():
    """TFX_PREC_AUTO on LoButterworth | ParametricEQ | HiShelving: float64 only in the EQ section
    (mixed-precision tile kernel); must meet the float64-recurrence tolerance, with state carry."""
    import torchfx_b200 as fx

    chain = [fx.filter.LoButterworth(5000, order=4, fs=48000), fx.filter.ParametricEQ(1000, q=2.0, gain=3.0, fs=48000),
             fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=48000)]
    for f in chain:
        f.compute_coefficients()
    sos = np.vstack([f._sos.numpy() for f in chain])
    rng = np.random.default_rng(4)
    x = (0.1 * rng.standard_normal((64, 400001))).astype(np.float32)[:, :400000]
    sx0 = 0.1 * rng.standard_normal((4, 64, 2))
    sy0 = 0.1 * rng.standard_normal((4, 64, 2))
    want, wsx, wsy = oracle.sos_cascade(x, sos, sx0, sy0)
    y, sx, sy, _ = run(x, sos, sx0.copy(), sy0.copy(), precision="auto")
    assert rel_to_max(y, want) < 2e-6
    np.testing.assert_allclose(sx, wsx, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sy, wsy, rtol=1e-4, atol=1e-5 * np.abs(wsy).max())
    # chunked == contiguous
    xt = torch.from_numpy(x).to(DEV)
    stx = torch.from_numpy(sx0).to(DEV)
    sty = torch.from_numpy(sy0).to(DEV)
    parts = [_ops.sos_cascade_(xt[:, lo:hi], torch.from_numpy(sos), stx, sty) for lo, hi in ((0, 100000), (100000, 100001), (100001, 400000))]
    assert rel_to_max(torch.cat(parts, 1).cpu().numpy(), want) < 2e-6


@pytest.mark.parametrize("where", ["first", "middle", "two"])
def test_mixed_precision_long_chains(where):
    """Cascades of 5..8 sections with float32-hostile sections (a 20 Hz high-pass, a narrow low notch): TFX_PREC_AUTO keeps
    float64 only where it is needed -- any single section or the first two are instantiated (sos_tile_mixed.cu), any other
    mask runs the float64 kernel -- and must meet the float64-recurrence tolerance, state included."""
    import torchfx_b200 as fx
    from torchfx_b200 import _native

    lo = fx.filter.LoButterworth(6000, order=6, fs=48000)      # 3 sections, float32-friendly
    shelf = fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=48000)
    hp = fx.filter.HiButterworth(20, order=2, fs=48000)        # float32-hostile
    notch = fx.filter.Notch(60, q=30.0, fs=48000) if hasattr(fx.filter, "Notch") else fx.filter.HiButterworth(30, order=2, fs=48000)
    eq = fx.filter.ParametricEQ(9000, q=1.0, gain=-2.0, fs=48000)
    chain = {"first": [hp, lo, shelf, eq], "middle": [lo, hp, shelf, eq], "two": [lo, hp, shelf, notch, eq]}[where]
    for f in chain:
        f.compute_coefficients()
    sos = np.vstack([f._sos.numpy() for f in chain])
    K = sos.shape[0]
    assert 5 <= K <= 8
    lib = _native.load()
    import ctypes

    err = ctypes.c_double(0.0)
    mask = lib.tfx_sos_mixed_mask(np.ascontiguousarray(sos).ctypes.data, K, ctypes.byref(err))
    assert 0 < mask < (1 << K) - 1, f"expected a proper float64 subset, got mask {mask:#x}"
    rng = np.random.default_rng(11)
    x = (0.1 * rng.standard_normal((64, 300000))).astype(np.float32)
    sx0 = 0.1 * rng.standard_normal((K, 64, 2))
    sy0 = 0.1 * rng.standard_normal((K, 64, 2))
    want, wsx, wsy = oracle.sos_cascade(x, sos, sx0, sy0)
    y, sx, sy, _ = run(x, sos, sx0.copy(), sy0.copy(), precision="auto")
    assert rel_to_max(y, want) < 2e-6
    np.testing.assert_allclose(sx, wsx, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sy, wsy, rtol=1e-4, atol=1e-5 * np.abs(wsy).max())
    y64, _, _, _ = run(x, sos, sx0.copy(), sy0.copy(), precision="f64")
    assert rel_to_max(y, y64) < 2e-6
