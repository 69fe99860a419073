"""GPU parity of the band-per-lane filterbank kernel (tfx_filterbank_*), both parallel
semantics of the reference: stacked (LogFilterBank, filter/filterbank.py:183-185) and summed
(ParallelFilterCombination, filter/__base.py:1019-1026).  Mirrors the reference's
tests/test_filter_base.py:145-203, tests/test_filterbank.py, tests/test_cuda_fallback.py:90-95."""
from __future__ import annotations

import numpy as np
import pytest
import scipy.signal as sps
import torch

import torchfx_b200 as fx
from conftest import golden, rel_to_max
from oracle import oracle
from torchfx_b200 import _native

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def test_logfilterbank_golden_stack():
    g = golden("parallel.npz")
    bank = fx.filter.LogFilterBank(n_bands=8, f_min=100.0, f_max=8000.0, fs=48000)
    before = _native.kernel_launches()
    y = bank(torch.from_numpy(g["x"]).to(DEV))
    torch.cuda.synchronize()
    assert 1 <= _native.kernel_launches() - before <= 4  # fused launches (f32 group, f64 group, optional warm-ups), not 8
    assert y.shape == (8, 2, 4096)
    assert rel_to_max(y.cpu().numpy(), g["bank"]) < TOL
    np.testing.assert_allclose(np.stack([f._sos.numpy() for f in bank.filters]), g["bank_sos"], rtol=1e-13)


def test_parallel_sum_golden():
    g = golden("parallel.npz")
    comb = fx.filter.BiquadBPF(500, 1.414, 48000) + fx.filter.BiquadBPF(2000, 1.414, 48000) + fx.filter.LoButterworth(300, order=2, fs=48000)
    assert len(comb.filters) == 2  # (a + b) + c nests like the reference
    y = comb(torch.from_numpy(g["x"]).to(DEV))
    assert rel_to_max(y.cpu().numpy(), g["comb"]) < TOL


@pytest.mark.parametrize("n_bands", [2, 3, 8, 17, 32, 40])
@pytest.mark.parametrize("mode", ["stack", "sum"])
def test_band_counts_vs_oracle(n_bands, mode):
    from torchfx_b200.filter._sosbank import SosBank

    rng = np.random.default_rng(n_bands)
    x = (0.1 * rng.standard_normal((5, 30000))).astype(np.float32)
    filters = [fx.filter.BiquadBPF(200.0 * (1.12 ** i), 1.414, 48000) for i in range(n_bands)]
    bank = SosBank(filters, mode=mode)
    y = bank(torch.from_numpy(x).to(DEV)).cpu().numpy()
    sos = np.stack([f._sos.numpy() for f in filters])
    want = oracle.filterbank_stack(x, sos) if mode == "stack" else oracle.filterbank_sum(x, sos)
    assert y.shape == want.shape
    assert rel_to_max(y, want) < TOL
    # every child's DF1 state is what running it alone would have left
    _, wsx, wsy = oracle.sos_cascade(x, sos[n_bands // 2])
    np.testing.assert_allclose(filters[n_bands // 2]._state_y.cpu().numpy(), wsy, rtol=1e-3, atol=1e-5 * np.abs(wsy).max())
    np.testing.assert_allclose(filters[n_bands // 2]._state_x.cpu().numpy(), wsx, rtol=1e-6, atol=1e-7)


def test_multi_section_children_and_chunked_state():
    """Children with different orders are padded with pass-through sections; state carries
    across calls exactly as with separately-run children."""
    rng = np.random.default_rng(99)
    x = (0.1 * rng.standard_normal((3, 50001))).astype(np.float32)  # odd length: unaligned rows
    def make():
        return [fx.filter.LoButterworth(3000, order=4, fs=48000), fx.filter.HiButterworth(300, order=6, fs=48000),
                fx.filter.ParametricEQ(1000, 2.0, 3.0, fs=48000)]
    comb = fx.filter._base.ParallelFilterCombination(*make())
    xt = torch.from_numpy(x).to(DEV)
    y = torch.cat([comb(xt[:, :20000]), comb(xt[:, 20000:])], dim=1).cpu().numpy()
    want = np.zeros_like(x)
    for f in make():
        f.compute_coefficients()
        want += oracle.sos_cascade(x, f._sos.numpy())[0]
    assert rel_to_max(y, want) < TOL


def test_low_band_forces_f64_and_long_signal_splits():
    """20 Hz band: long time constant (warm-up ~3e4 samples) and float32-hostile."""
    rng = np.random.default_rng(5)
    x = (0.1 * rng.standard_normal((2, 1 << 20))).astype(np.float32)
    bank = fx.filter.LogFilterBank(n_bands=4, f_min=20.0, f_max=2000.0, fs=48000)
    y = bank(torch.from_numpy(x).to(DEV)).cpu().numpy()
    bank.compute_coefficients()
    sos = np.stack([f._sos.numpy() for f in bank.filters])
    want = oracle.filterbank_stack(x, sos)
    assert rel_to_max(y, want) < 1e-6


def test_f64_io():
    rng = np.random.default_rng(8)
    x = rng.standard_normal((2, 9000))
    bank = fx.filter.LogFilterBank(n_bands=5, f_min=100.0, f_max=5000.0, fs=44100)
    y = bank(torch.from_numpy(x).to(DEV))
    assert y.dtype == torch.float64
    bank.compute_coefficients()
    want = oracle.filterbank_stack(x, np.stack([f._sos.numpy() for f in bank.filters]))
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-9, atol=1e-11)
