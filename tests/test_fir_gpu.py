"""GPU parity of the FIR kernels (tfx_fir_f32: shared-memory direct form and the persistent partitioned
overlap-save kernel with its own 16384-point shared-memory FFTs) against the f64-accumulating oracle and the
reference-generated golden vectors.  Mirrors the reference's tests/test_fftconv.py:65-123 and
tests/test_fir.py:14-131 (their tolerance: 1e-4; here 1e-5 of max|y| as north_star asks)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import torchfx_b200 as fx
from conftest import golden, rel_to_max
from oracle import oracle
from torchfx_b200 import _native
from torchfx_b200.filter._fftconv import fft_conv1d
from torchfx_b200.filter.fir import fir_causal, fir_plan

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def test_golden_fft_and_direct_modes():
    g = golden("fir.npz")
    x = torch.from_numpy(g["x"]).to(DEV)
    before = _native.kernel_launches()
    y_fft = fx.filter.FIR(g["taps"])(x)
    y_dir = fx.filter.FIR(g["taps"], conv_mode="direct")(x)
    assert _native.kernel_launches() > before
    assert rel_to_max(y_fft.cpu().numpy(), g["y_fft"]) < TOL
    assert rel_to_max(y_dir.cpu().numpy(), g["y_direct"]) < TOL
    f = fx.filter.FIR(g["taps"])
    assert f.kernel.shape == (1, 1, 101) and f.kernel.dtype == torch.float32
    np.testing.assert_array_equal(f.kernel[0, 0].numpy(), g["taps"][::-1])  # stored flipped, like the reference


def test_designable_fir_golden():
    g = golden("fir.npz")
    des = fx.filter.DesignableFIR(cutoff=3000.0, num_taps=63, fs=48000)
    np.testing.assert_allclose(np.asarray(des.b), g["des_b"], rtol=1e-12)
    y = des(torch.from_numpy(g["x"]).to(DEV))
    assert rel_to_max(y.cpu().numpy(), g["y_des"]) < TOL


def test_fft_conv1d_golden_and_errors():
    g = golden("fir.npz")
    xc = torch.from_numpy(g["xc"]).to(DEV)
    kern = torch.from_numpy(g["kern"]).to(DEV)
    y = fft_conv1d(xc, kern)
    assert y.shape == g["yc"].shape and rel_to_max(y.cpu().numpy(), g["yc"]) < TOL
    yp = fft_conv1d(xc, kern, padding=(5, 10))
    assert yp.shape == g["yc_pad"].shape and rel_to_max(yp.cpu().numpy(), g["yc_pad"]) < TOL
    with pytest.raises(RuntimeError, match="at least as large"):
        fft_conv1d(xc[..., :10], kern)
    with pytest.raises(RuntimeError, match="Block ratio"):
        fft_conv1d(xc, kern, block_ratio=0.5)


@pytest.mark.parametrize("algo", [_native.TFX_FIR_DIRECT, _native.TFX_FIR_OLS])
@pytest.mark.parametrize("K", [1, 2, 7, 8, 9, 64, 97, 256, 1000])
def test_tap_counts_both_algorithms(algo, K):
    rng = np.random.default_rng(K)
    x = rng.standard_normal((3, 10007)).astype(np.float32)
    b = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    want = oracle.fir_causal(x, b)
    y = fir_causal(torch.from_numpy(x).to(DEV), torch.from_numpy(b), algo)
    assert rel_to_max(y.cpu().numpy(), want) < TOL


@pytest.mark.parametrize("K,T", [(2048, 5000), (2049, 5000), (5000, 3000), (4096, 4096), (12345, 40000), (300, 100),
                                 (20000, 30000), (70000, 150000), (131072, 140000), (8192, 8192), (8193, 16385), (16384, 100000)])
def test_partition_edges(K, T):
    """K around the partition sizes (2048 first version / 8192), K > T, T around a block; partition counts that are
    not a multiple of the 8-partition chunk (9, 16 partitions) and exactly one partition."""
    rng = np.random.default_rng(K + T)
    x = rng.standard_normal((2, T)).astype(np.float32)
    b = (rng.standard_normal(K) * np.exp(-np.arange(K) / (K / 4))).astype(np.float32)
    want = oracle.fir_causal(x, b)
    y = fir_causal(torch.from_numpy(x).to(DEV), torch.from_numpy(b), _native.TFX_FIR_OLS)
    assert rel_to_max(y.cpu().numpy(), want) < TOL


@pytest.mark.parametrize("K", [300, 8192, 20000, 65536])
def test_plan_is_bit_identical_and_cached(K):
    """tfx_fir_plan_init + tfx_fir_f32_planned (twiddles and taps spectra computed once) against tfx_fir_f32 with
    TFX_FIR_OLS, bit for bit; the FIR module builds the plan on its first call and reuses it for later chunks, and a
    changed kernel buffer (in-place edit bumps the tensor version) invalidates it."""
    rng = np.random.default_rng(K)
    b = (rng.standard_normal(K) * np.exp(-np.arange(K) / (K / 5.0))).astype(np.float32)
    x = torch.from_numpy(rng.standard_normal((6, 50000)).astype(np.float32)).to(DEV)
    bt = torch.from_numpy(b).to(DEV)
    y0 = fir_causal(x, bt, _native.TFX_FIR_OLS)
    plan = fir_plan(bt)
    y1 = fir_causal(x, bt, _native.TFX_FIR_OLS, plan=plan)
    assert torch.equal(y0, y1)
    f = fx.filter.FIR(b).to(DEV)
    n0 = _native.load().tfx_kernel_launches()
    ya = f(x)
    n1 = _native.load().tfx_kernel_launches()
    yb = f(x[:, :30000])
    n2 = _native.load().tfx_kernel_launches()
    if K > 96:  # overlap-save (AUTO takes the direct form below 97 taps)
        assert n1 - n0 == 3 and n2 - n1 == 1, (n1 - n0, n2 - n1)  # plan (2 kernels) + filter, then the filter alone
    assert torch.equal(ya, y0) and torch.equal(yb, fir_causal(x[:, :30000], bt, _native.TFX_FIR_OLS))
    f.kernel.mul_(2.0)
    yc = f(x)
    assert rel_to_max(yc.cpu().numpy(), 2.0 * y0.cpu().numpy()) < 1e-6
    f.kernel = f.kernel * 0.25  # a REPLACED buffer (new tensor object, version 0): the plan must follow it too
    yd = f(x)
    assert rel_to_max(yd.cpu().numpy(), 0.5 * y0.cpu().numpy()) < 1e-6


def test_cfg3_reverb_ir_65536_taps():
    """BASELINE configs[2] at oracle-checkable size: 65 536-tap decaying-noise IR,
    5 channels (odd: one half-empty pair) x 100 000 samples."""
    rng = np.random.default_rng(7)
    K = 65536
    ir = rng.standard_normal(K) * np.exp(-np.arange(K) / 8000.0)
    ir = (ir / np.sqrt((ir ** 2).sum())).astype(np.float32)
    x = (0.1 * rng.standard_normal((5, 100000))).astype(np.float32)
    y = fx.filter.FIR(ir)(torch.from_numpy(x).to(DEV)).cpu().numpy()
    want = oracle.fir_causal(x, ir)
    assert rel_to_max(y, want) < TOL


@pytest.mark.parametrize("K", [5000, 40000])
def test_many_channels_multiple_slabs_linearity(K):
    """256 channels x 600k samples: the spectra workspace is processed in several time slabs and the
    spectra ring wraps (3 and 20 partitions of history carried from slab to slab).
    Checked by linearity + a 4-channel oracle comparison (the full oracle would take minutes)."""
    g = torch.Generator(device=DEV).manual_seed(5)
    x1 = 0.1 * torch.randn(256, 600_000, device=DEV, generator=g)
    x2 = 0.1 * torch.randn(256, 600_000, device=DEV, generator=g)
    rng = np.random.default_rng(11)
    b = (rng.standard_normal(K) * np.exp(-np.arange(K) / (K / 7.0))).astype(np.float32)
    bt = torch.from_numpy(b)
    y1 = fir_causal(x1, bt)
    y2 = fir_causal(x2, bt)
    y12 = fir_causal(x1 + 2.0 * x2, bt)
    assert float((y12 - (y1 + 2.0 * y2)).abs().max() / y12.abs().max()) < 2e-5
    pick = [0, 1, 128, 255]
    want = oracle.fir_causal(x1[pick].cpu().numpy(), b)
    assert rel_to_max(y1[pick].cpu().numpy(), want) < TOL


def test_shapes_and_dtype_like_reference():
    b = np.hanning(33).astype(np.float32)
    f = fx.filter.FIR(b)
    for shape in [(1000,), (2, 1000), (3, 2, 1000)]:
        x = torch.randn(*shape, device=DEV)
        y = f(x)
        assert y.shape == x.shape and y.dtype == x.dtype and y.is_cuda
    with pytest.raises(ValueError, match="conv_mode"):
        fx.filter.FIR(b, conv_mode="nope")


@pytest.mark.parametrize("knobs", [{}, {"TFX_FIR_G": "4", "TFX_FIR_LM": "1", "TFX_FIR_LI": "2"}, {"TFX_FIR_G": "12", "TFX_FIR_LM": "5", "TFX_FIR_LI": "10"}])
@pytest.mark.parametrize("K,T,C", [(3000, 20000, 3), (9000, 50001, 2), (33000, 70000, 5), (65536, 40000, 1), (100000, 300000, 19)])
def test_queue_settings(knobs, K, T, C, monkeypatch):
    """The persistent overlap-save kernel under different queue settings (open channel-pair slots G, queue lags of the
    multiply and inverse items: the schedule changes, the result must not); odd channel counts leave a half-empty pair, T is not a multiple of the 8192-sample hop, K is not a
    multiple of the partition, 19 channels need more than one group of slots."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(K + T)
    x = rng.standard_normal((C, T)).astype(np.float32)
    b = (rng.standard_normal(K) * np.exp(-np.arange(K) / (K / 5.0))).astype(np.float32)
    want = oracle.fir_causal(x, b)
    y = fir_causal(torch.from_numpy(x).to(DEV), torch.from_numpy(b), _native.TFX_FIR_OLS)
    assert rel_to_max(y.cpu().numpy(), want) < TOL


def test_unaligned_rows_and_views():
    """Row starts that are not 8-byte aligned (odd leading dimension / offset view) take the scalar load / store path."""
    rng = np.random.default_rng(3)
    base = torch.from_numpy(rng.standard_normal((3, 50001)).astype(np.float32)).to(DEV)
    x = base[:, 1:]  # rows start at odd element offsets
    b = (rng.standard_normal(9000) * np.exp(-np.arange(9000) / 2000.0)).astype(np.float32)
    y = fir_causal(x, torch.from_numpy(b), _native.TFX_FIR_OLS)
    want = oracle.fir_causal(x.cpu().numpy(), b)
    assert rel_to_max(y.cpu().numpy(), want) < TOL


@pytest.mark.parametrize("K,T,C", [(1, 100, 1), (33, 5000, 3), (512, 4000, 2), (513, 2049, 2), (3000, 10000, 2)])
def test_float64_signals_are_filtered_in_float64(K, T, C):
    """The reference evaluates FIR in the input dtype (filter/fir.py:529-531): a float64 CUDA signal runs the float64
    direct-form kernel (tfx_fir_f64) and must match the float64 oracle to float64 accuracy, not float32's."""
    rng = np.random.default_rng(K + T)
    x = rng.standard_normal((C, T))
    b = rng.standard_normal(K) / np.sqrt(K)
    import scipy.signal as sps

    want = sps.lfilter(b, [1.0], x, axis=-1)  # float64 direct form, the oracle's FIR entry is float32-I/O
    y = fir_causal(torch.from_numpy(x).to(DEV), torch.from_numpy(b))
    assert y.dtype == torch.float64
    assert rel_to_max(y.cpu().numpy(), want) < 1e-12
    f = fx.filter.FIR(b.astype(np.float32))  # the module stores float32 taps, like the reference (fir.py:516-518)
    y2 = f(torch.from_numpy(x).to(DEV))
    assert y2.dtype == torch.float64
    assert rel_to_max(y2.cpu().numpy(), sps.lfilter(b.astype(np.float32).astype(np.float64), [1.0], x, axis=-1)) < 1e-12


@pytest.mark.parametrize("K", [33, 97, 1024, 1025, 1026, 2048, 2049, 5000, 7168, 7169, 8191, 8192])
@pytest.mark.parametrize("C,T", [(3, 2 * 15360 + 7), (2, 40001), (1, 9000)])
def test_single_partition_hop_boundaries(K, C, T):
    """K <= 8192 runs without a frequency-domain delay line, and its blocks hop by the largest multiple of 1024 that is
    <= 16384 - K + 1 (15 360 samples up to 1025 taps, 14 336 from 1026, ..., 8192 from 7170): tap counts on both sides of
    every kind of boundary, signals that end inside a block, one that is shorter than a block, an odd channel count."""
    rng = np.random.default_rng(K * 7 + T)
    x = rng.standard_normal((C, T)).astype(np.float32)
    b = (rng.standard_normal(K) * np.exp(-np.arange(K) / (K / 4.0))).astype(np.float32)
    want = oracle.fir_causal(x, b)
    y = fir_causal(torch.from_numpy(x).to(DEV), torch.from_numpy(b), _native.TFX_FIR_OLS)
    assert rel_to_max(y.cpu().numpy(), want) < TOL
