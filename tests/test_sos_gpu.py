"""GPU parity of the fused SOS cascade, through the C ABI (tfx_sos_cascade_*), against the
CPU oracle and the reference-generated golden vectors.

Tolerance (BASELINE.json north_star): 1e-5 of max|y_ref| per channel for float32 I/O.
The float64 recurrence is held to 1e-6 (it should be ~1e-7: one float32 rounding), float64
I/O to 1e-10.  Mirrors the reference's tests/test_cuda_kernels.py:35-184,
tests/test_fused.py:116-259 and tests/test_ops_dispatch.py:48-127.
"""
from __future__ import annotations

import numpy as np
import pytest
import scipy.signal as sps
import torch

import torchfx_b200 as fx
from conftest import golden, rel_to_max
from oracle import oracle
from torchfx_b200 import _native, _ops

pytestmark = pytest.mark.gpu
TOL_F32 = 1e-5
TOL_F64REC = 1e-6
DEV = "cuda:0"


def run(x_np, sos_np, sx=None, sy=None, precision="auto", no_split=False, inplace=False):
    x = torch.from_numpy(np.ascontiguousarray(x_np)).to(DEV)
    K, C = sos_np.shape[0], x.shape[0]
    stx = torch.zeros(K, C, 2, dtype=torch.float64, device=DEV) if sx is None else torch.from_numpy(sx).to(DEV)
    sty = torch.zeros(K, C, 2, dtype=torch.float64, device=DEV) if sy is None else torch.from_numpy(sy).to(DEV)
    before = _native.kernel_launches()
    y = _ops.sos_cascade_(x, torch.from_numpy(sos_np), stx, sty, out=x if inplace else None, precision=precision, no_split=no_split)
    torch.cuda.synchronize()
    assert _native.kernel_launches() > before, "the CUDA path did not launch a kernel"
    return y.cpu().numpy(), stx.cpu().numpy(), sty.cpu().numpy()


def test_cfg1_golden():
    g = golden("cfg1_lobutter4_mono.npz")
    y, _, _ = run(g["x"], g["sos"])
    assert y.dtype == np.float32
    assert rel_to_max(y, g["y"]) < TOL_F32


@pytest.mark.parametrize("precision,tol", [("f32", TOL_F32), ("f64", TOL_F64REC), ("auto", TOL_F32)])
def test_k4_golden_output_and_state(precision, tol):
    g = golden("sos_k4_chunked.npz")
    y, sx, sy = run(g["x"], g["sos"], precision=precision)
    assert rel_to_max(y, g["y"]) < tol
    np.testing.assert_allclose(sx, g["state_x"], rtol=tol * 10, atol=tol * np.abs(g["state_x"]).max())
    np.testing.assert_allclose(sy, g["state_y"], rtol=tol * 10, atol=tol * np.abs(g["state_y"]).max())


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_chunked_equals_contiguous(precision):
    g = golden("sos_k4_chunked.npz")
    split = int(g["split"])
    ya, sxa, sya = run(g["x"][:, :split], g["sos"], precision=precision)
    np.testing.assert_allclose(sxa, g["state_x_mid"], rtol=1e-4, atol=1e-6)
    yb, sxb, syb = run(g["x"][:, split:], g["sos"], sxa, sya, precision=precision)
    assert rel_to_max(np.concatenate([ya, yb], 1), g["y"]) < TOL_F32
    np.testing.assert_allclose(syb, g["state_y"], rtol=1e-3, atol=1e-5 * np.abs(g["state_y"]).max())


def test_f64_io_with_initial_state_golden():
    g = golden("ops_state_f64.npz")
    y, sx, sy = run(g["x"], g["sos"], g["sx0"], g["sy0"])
    np.testing.assert_allclose(y, g["y"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(sx, g["sx1"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(sy, g["sy1"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6, 7, 8, 10, 16])
def test_every_section_count_vs_oracle(K):
    rng = np.random.default_rng(K)
    x = (0.1 * rng.standard_normal((5, 20000))).astype(np.float32)
    sos = sps.butter(2 * K, 0.21, output="sos")
    want, wsx, wsy = oracle.sos_cascade(x, sos)
    for precision, tol in (("f32", TOL_F32), ("f64", TOL_F64REC)):
        y, sx, sy = run(x, sos, precision=precision)
        assert rel_to_max(y, want) < tol, (K, precision)
        np.testing.assert_allclose(sy, wsy, rtol=1e-3, atol=tol * 10 * max(np.abs(wsy).max(), 1e-30))


@pytest.mark.parametrize("shape", [(1, 1), (1, 2), (3, 3), (2, 63), (2, 64), (2, 65), (7, 257), (33, 1000), (2, 4099), (130, 300)])
def test_ragged_and_tiny_shapes(shape):
    """Odd lengths (unaligned rows -> element-wise path), T < 2 (state hand-over), C not a
    multiple of 32 (dead lanes)."""
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    sos = sps.cheby1(4, 1.0, 0.3, output="sos")
    sx0 = rng.standard_normal((2, shape[0], 2))
    sy0 = rng.standard_normal((2, shape[0], 2))
    want, wsx, wsy = oracle.sos_cascade(x, sos, sx0, sy0)
    y, sx, sy = run(x, sos, sx0.copy(), sy0.copy(), precision="f64")
    assert rel_to_max(y, want) < TOL_F64REC
    np.testing.assert_allclose(sx, wsx, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(sy, wsy, rtol=1e-5, atol=1e-5)


def test_empty_input_is_a_noop():
    sos = torch.from_numpy(sps.butter(2, 0.2, output="sos"))
    st = torch.ones(1, 3, 2, dtype=torch.float64, device=DEV)
    y = _ops.sos_cascade_(torch.zeros(3, 0, device=DEV), sos, st, st.clone())
    assert y.shape == (3, 0) and torch.all(st == 1)


def test_time_split_equals_single_stream():
    """Few channels x long signal: the planner cuts every channel into many segments with
    warm-up; the result must equal the unsplit sequential run (TFX_NO_SPLIT) and the oracle."""
    rng = np.random.default_rng(3)
    x = (0.1 * rng.standard_normal((2, 1 << 20))).astype(np.float32)
    sos = sps.butter(8, 5000 / 24000, output="sos")
    want, _, wsy = oracle.sos_cascade(x, sos)
    for precision, tol in (("f32", TOL_F32), ("f64", TOL_F64REC)):
        ys, _, sys_ = run(x, sos, precision=precision)
        yn, _, syn = run(x, sos, precision=precision, no_split=True)
        assert rel_to_max(ys, want) < tol
        assert rel_to_max(yn, want) < tol
        # split vs unsplit differ only by the warm-up truncation (<= 2^-30 of the state)
        assert rel_to_max(ys, yn) < 2e-6
        np.testing.assert_allclose(sys_, syn, rtol=1e-4, atol=1e-6 * np.abs(wsy).max())


def test_low_corner_filter_takes_f64_and_stays_in_tolerance():
    """20 Hz high-pass: hopeless in float32 (SURVEY.md 7 hard part 1: 4e-4), must be routed
    to the float64 recurrence by TFX_PREC_AUTO."""
    rng = np.random.default_rng(5)
    x = (0.1 * rng.standard_normal((4, 400000))).astype(np.float32)
    sos = sps.butter(2, 20 / 24000, btype="highpass", output="sos")
    want, _, _ = oracle.sos_cascade(x, sos)
    y, _, _ = run(x, sos, precision="auto")
    assert rel_to_max(y, want) < TOL_F64REC


def test_in_place():
    rng = np.random.default_rng(11)
    x = (0.1 * rng.standard_normal((3, 300000))).astype(np.float32)
    sos = sps.butter(8, 5000 / 24000, output="sos")
    want, _, _ = oracle.sos_cascade(x, sos)
    y, _, _ = run(x, sos, inplace=True)
    assert rel_to_max(y, want) < TOL_F32


def test_unstable_filter_is_never_split_and_matches():
    """A pole outside the unit circle: no decay bound exists -> one stream per channel."""
    rng = np.random.default_rng(13)
    x = rng.standard_normal((2, 300)).astype(np.float64)
    sos = np.array([[1.0, 0.0, 0.0, 1.0, -2.02, 1.0201]])  # double pole at 1.01
    want, _, _ = oracle.sos_cascade(x, sos)
    y, _, _ = run(x, sos)
    np.testing.assert_allclose(y, want, rtol=1e-9)


def test_noncontiguous_rows_and_row_stride():
    rng = np.random.default_rng(17)
    big = torch.from_numpy((0.1 * rng.standard_normal((6, 5000))).astype(np.float32)).to(DEV)
    view = big[1:5, 8:4008]  # row stride 5000, 32-byte aligned start
    sos = sps.butter(4, 0.2, output="sos")
    want, _, _ = oracle.sos_cascade(view.cpu().numpy(), sos)
    y = _ops.sos_cascade_(view, torch.from_numpy(sos), None, None)
    assert rel_to_max(y.cpu().numpy(), want) < TOL_F32
    view2 = big[:, 3:4002]  # misaligned start -> element-wise path
    want2, _, _ = oracle.sos_cascade(view2.cpu().numpy(), sos)
    y2 = _ops.sos_cascade_(view2, torch.from_numpy(sos), None, None)
    assert rel_to_max(y2.cpu().numpy(), want2) < TOL_F32


def test_module_surface_on_cuda_matches_scipy():
    """Reference tests/test_cuda_kernels.py:35-65 with the move_coeff shim it expects."""
    torch.manual_seed(0)
    x = torch.randn(2, 44100, dtype=torch.float64)
    f = fx.filter.LoButterworth(1000, order=6, fs=44100)
    f.compute_coefficients()
    f.move_coeff("cuda")
    y = f(x.to(DEV))
    ref = sps.sosfilt(f._sos.numpy(), x.numpy(), axis=-1)
    np.testing.assert_allclose(y.cpu().numpy(), ref, atol=1e-4, rtol=1e-4)
    assert y.dtype == torch.float64 and y.is_cuda
    assert f._state_x.shape == (3, 2, 2) and f._state_x.is_cuda


def test_cfg4_fused_chain_golden_and_single_launch_pair():
    g = golden("cfg4_chain.npz")
    x = torch.from_numpy(g["x"])
    before = _native.kernel_launches()
    w = (fx.Wave(x, 48000, device=DEV) | fx.filter.LoButterworth(5000, order=4) | fx.filter.ParametricEQ(1000, q=2.0, gain=3.0)
         | fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db"))
    y = w.ys
    torch.cuda.synchronize()
    # three piped filters -> ONE fused cascade call (<= 2 launches: optional warm-up + main)
    assert 1 <= _native.kernel_launches() - before <= 2
    assert rel_to_max(y.cpu().numpy(), g["y"]) < TOL_F32


def test_linearity_and_sortedness_properties_at_scale():
    """Size-independent properties on a larger block (256 ch x 2^18): linearity of the
    filter and agreement of 8 random channels with the oracle over the full length."""
    g = torch.Generator(device=DEV).manual_seed(1234)
    x1 = 0.1 * torch.randn(256, 1 << 18, device=DEV, generator=g)
    x2 = 0.1 * torch.randn(256, 1 << 18, device=DEV, generator=g)
    sos_np = sps.butter(8, 5000 / 24000, output="sos")
    sos = torch.from_numpy(sos_np)
    y1 = _ops.sos_cascade_(x1, sos, None, None)
    y2 = _ops.sos_cascade_(x2, sos, None, None)
    y12 = _ops.sos_cascade_(x1 + 2.0 * x2, sos, None, None)
    lin = (y12 - (y1 + 2.0 * y2)).abs().max() / y12.abs().max()
    assert float(lin) < 2e-5
    pick = [0, 7, 31, 32, 100, 128, 200, 255]
    want, _, _ = oracle.sos_cascade(x1[pick].cpu().numpy(), sos_np)
    assert rel_to_max(y1[pick].cpu().numpy(), want) < TOL_F32


def test_delay_line_golden():
    g = golden("delay.npz")
    y = _ops.delay_line_forward(torch.from_numpy(g["x"]).to(DEV), 100, 0.5, 0.8)
    assert rel_to_max(y.cpu().numpy(), g["y"]) < 1e-6
    y64 = _ops.delay_line_forward(torch.from_numpy(g["x64"]).to(DEV), 333, 0.7, 0.25)
    np.testing.assert_allclose(y64.cpu().numpy(), g["y64"], rtol=1e-13, atol=1e-14)
    short = torch.randn(2, 50, device=DEV)
    assert _ops.delay_line_forward(short, 100, 0.5, 0.5) is short


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("C,T,delay,view", [(3, 10007, 100, False), (2, 4096, 4095, False), (5, 20001, 4410, False), (2, 9001, 7, True),
                                             (1, 12289, 8192, False), (4, 5000, 1, True)])
def test_delay_line_shapes_vs_oracle(dtype, C, T, delay, view):
    """The tiled delay-line kernel against the oracle (delay_cpu.cpp:17-85): delays that are / are not multiples of the
    vector width, longer than a 4096-sample tile, T not a multiple of the tile, rows that are not 16-byte aligned (view)."""
    rng = np.random.default_rng(C * T + delay)
    x = rng.standard_normal((C, T + 1)).astype(dtype)
    xt = torch.from_numpy(x).to(DEV)
    xin, xnp = (xt[:, 1:], x[:, 1:]) if view else (xt[:, :T], x[:, :T])
    y = _ops.delay_line_forward(xin, delay, 0.6, 0.75)
    want = oracle.delay_line(np.ascontiguousarray(xnp), delay, 0.6, 0.75)
    assert y.shape == want.shape and y.dtype == xin.dtype
    if dtype == np.float32:
        assert rel_to_max(y.cpu().numpy(), want) < 1e-6
    else:
        np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-13, atol=1e-14)


def test_host_streaming_driver_matches_oracle():
    import ctypes

    rng = np.random.default_rng(23)
    C, T = 16, 200000
    x = torch.from_numpy((0.1 * rng.standard_normal((C, T))).astype(np.float32)).pin_memory()
    y = torch.empty_like(x).pin_memory()
    sos_np = sps.butter(8, 5000 / 24000, output="sos")
    sos = torch.from_numpy(sos_np).contiguous()
    sx = torch.zeros(4, C, 2, dtype=torch.float64)
    sy = torch.zeros(4, C, 2, dtype=torch.float64)
    lib = _native.load()
    _native.check(lib.tfx_sos_cascade_host_f32(x.data_ptr(), y.data_ptr(), C, T, T, T, sos.data_ptr(), 4, sx.data_ptr(), sy.data_ptr(), 0, 30000, 0))
    want, wsx, wsy = oracle.sos_cascade(x.numpy(), sos_np)
    assert rel_to_max(y.numpy(), want) < TOL_F32
    np.testing.assert_allclose(sx.numpy(), wsx, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sy.numpy(), wsy, rtol=1e-3, atol=1e-5 * np.abs(wsy).max())


def test_unaligned_rows_are_repitched_and_match():
    """Odd T with many channels: the Python layer re-pitches once instead of taking the
    element-wise kernel path; result and state must be unchanged, input untouched."""
    rng = np.random.default_rng(41)
    x = (0.1 * rng.standard_normal((64, 100003))).astype(np.float32)
    sos = sps.butter(8, 5000 / 24000, output="sos")
    want, wsx, wsy = oracle.sos_cascade(x, sos)
    xt = torch.from_numpy(x).to(DEV)
    sx = torch.zeros(4, 64, 2, dtype=torch.float64, device=DEV)
    sy = torch.zeros_like(sx)
    y = _ops.sos_cascade_(xt, torch.from_numpy(sos), sx, sy)
    assert y.shape == (64, 100003) and torch.equal(xt.cpu(), torch.from_numpy(x))
    assert rel_to_max(y.cpu().numpy(), want) < TOL_F32
    np.testing.assert_allclose(sy.cpu().numpy(), wsy, rtol=1e-3, atol=1e-5 * np.abs(wsy).max())
    # explicit out= keeps the direct (element-wise) path
    out = torch.empty_like(xt)
    y2 = _ops.sos_cascade_(xt, torch.from_numpy(sos), None, None, out=out)
    assert y2 is out and rel_to_max(out.cpu().numpy(), want) < TOL_F32


def test_config2_full_size_parity_as_survey_8d():
    """BASELINE configs[1] at FULL size (1024 ch x 28.8 M samples = 118 GB, in place), the parity check SURVEY.md 8d
    prescribes: ALL 1024 channels on the first 2**20 samples + 8 seeded-random channels over the full 10 minutes
    (long-run drift of the segmented float32 recurrence), both against the CPU oracle.  Needs a 180 GB GPU."""
    import gc

    gc.collect()
    torch.cuda.empty_cache()  # earlier tests leave tens of GB in torch's caching allocator
    free, _ = torch.cuda.mem_get_info()
    C, T, HEAD = 1024, 28_800_000, 1 << 20
    if free < C * T * 4 + (8 << 30):
        pytest.skip("not enough free HBM for the full-size configuration")
    g = torch.Generator(device=DEV).manual_seed(1234)
    x = torch.empty((C, T), device=DEV)
    x.normal_(0.0, 0.1, generator=g)
    pick = sorted(np.random.default_rng(8).choice(C, size=8, replace=False).tolist())
    x_pick = x[pick].cpu().numpy()
    x_head = x[:, :HEAD].cpu().numpy()
    sos_np = sps.butter(8, 5000 / 24000, output="sos")
    before = _native.kernel_launches()
    _ops.sos_cascade_(x, torch.from_numpy(sos_np), None, None, out=x)
    torch.cuda.synchronize()
    assert _native.kernel_launches() - before == 2
    want_pick, _, _ = oracle.sos_cascade(x_pick, sos_np)
    assert rel_to_max(x[pick].cpu().numpy(), want_pick) < TOL_F32
    got_head = x[:, :HEAD].cpu().numpy()
    del x
    torch.cuda.empty_cache()
    want_head, _, _ = oracle.sos_cascade(x_head, sos_np)  # causal: the head of y depends on the head of x only
    assert rel_to_max(got_head, want_head) < TOL_F32


def test_config2_scale_long_run_drift():
    """BASELINE configs[1] at a quarter of its length (1024 ch x 7.2 M samples, 29.5 GB, in place):
    8 seeded-random channels are checked against the oracle over the FULL length (long-run
    drift of the segmented float32 recurrence), plus the final DF1 state of those channels."""
    free, _ = torch.cuda.mem_get_info()
    C, T = 1024, 7_200_000
    if free < C * T * 4 * 1.2:
        pytest.skip("not enough free HBM")
    g = torch.Generator(device=DEV).manual_seed(1234)
    x = torch.empty((C, T), device=DEV)
    x.normal_(0.0, 0.1, generator=g)
    pick = sorted(np.random.default_rng(2).choice(C, size=8, replace=False).tolist())
    x_pick = x[pick].cpu().numpy()
    sos_np = sps.butter(8, 5000 / 24000, output="sos")
    sx = torch.zeros(4, C, 2, dtype=torch.float64, device=DEV)
    sy = torch.zeros_like(sx)
    before = _native.kernel_launches()
    _ops.sos_cascade_(x, torch.from_numpy(sos_np), sx, sy, out=x)
    torch.cuda.synchronize()
    assert _native.kernel_launches() - before == 2
    want, wsx, wsy = oracle.sos_cascade(x_pick, sos_np)
    assert rel_to_max(x[pick].cpu().numpy(), want) < TOL_F32
    np.testing.assert_allclose(sy[:, pick].cpu().numpy(), wsy, rtol=1e-3, atol=1e-5 * np.abs(wsy).max())
    np.testing.assert_allclose(sx[:, pick].cpu().numpy(), wsx, rtol=1e-3, atol=1e-5 * np.abs(wsx).max())
    del x
    torch.cuda.empty_cache()
