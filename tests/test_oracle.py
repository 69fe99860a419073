"""Pins the CPU oracle (oracle/) -- against golden vectors produced by the unmodified
reference (oracle/make_golden.py) and against scipy.signal, the oracle the reference's own
tests use (tests/test_ops_dispatch.py:74-127, tests/test_fused.py:116-137,
tests/test_fftconv.py:65-77).  CPU only."""
from __future__ import annotations

import numpy as np
import pytest
import scipy.signal as sps

from conftest import golden, rel_to_max
from oracle import oracle


def test_cfg1_matches_reference_golden():
    g = golden("cfg1_lobutter4_mono.npz")
    y, _, _ = oracle.sos_cascade(g["x"], g["sos"])
    assert y.dtype == np.float32
    # reference is built with -ffast-math: equal to a float32 ulp, not necessarily bitwise
    assert rel_to_max(y, g["y"]) < 2e-7
    np.testing.assert_allclose(g["sos"], sps.butter(4, 5000 / 24000, output="sos"), rtol=0, atol=0)


def test_k4_chunked_state_matches_reference_golden():
    g = golden("sos_k4_chunked.npz")
    y, sx, sy = oracle.sos_cascade(g["x"], g["sos"])
    assert rel_to_max(y, g["y"]) < 2e-7
    np.testing.assert_allclose(sx, g["state_x"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(sy, g["state_y"], rtol=1e-9, atol=1e-13)
    split = int(g["split"])
    ya, sxa, sya = oracle.sos_cascade(g["x"][:, :split], g["sos"])
    np.testing.assert_allclose(sxa, g["state_x_mid"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(sya, g["state_y_mid"], rtol=1e-9, atol=1e-13)
    yb, sxb, syb = oracle.sos_cascade(g["x"][:, split:], g["sos"], sxa, sya)
    assert np.array_equal(np.concatenate([ya, yb], axis=1), y)  # chunking is exact in the oracle
    np.testing.assert_array_equal(sxb, sx)


def test_ops_with_initial_state_f64_matches_reference_golden():
    g = golden("ops_state_f64.npz")
    y, sx, sy = oracle.sos_cascade(g["x"], g["sos"], g["sx0"], g["sy0"])
    np.testing.assert_allclose(y, g["y"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(sx, g["sx1"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(sy, g["sy1"], rtol=1e-10, atol=1e-12)
    sos = g["sos"]
    yb, bsx, bsy = oracle.biquad(g["x"], sos[0, :3], sos[0, 4], sos[0, 5], g["sx0"][0], g["sy0"][0])
    np.testing.assert_allclose(yb, g["yb"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(bsx, g["bsx"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(bsy, g["bsy"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("seed", [0, 1, 2, 42])
@pytest.mark.parametrize("order", [2, 4, 8, 12])
def test_sos_matches_scipy_sosfilt(seed, order):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((2, 6000))
    sos = sps.butter(order, 2000 / 22050, output="sos")
    y, _, _ = oracle.sos_cascade(x, sos)
    np.testing.assert_allclose(y, sps.sosfilt(sos, x, axis=-1), rtol=1e-9, atol=1e-11)


def test_cfg4_chain_matches_reference_golden():
    g = golden("cfg4_chain.npz")
    y, _, _ = oracle.sos_cascade(g["x"], g["sos"])
    assert rel_to_max(y, g["y"]) < 2e-7


def test_parallel_semantics_match_reference_golden():
    g = golden("parallel.npz")
    assert rel_to_max(oracle.filterbank_stack(g["x"], g["bank_sos"]), g["bank"]) < 2e-7
    lp = sps.butter(2, 300 / 24000, output="sos")

    def bpf(fc, q, fs=48000):
        w0 = 2 * np.pi * fc / fs
        al = np.sin(w0) / (2 * q)
        inv = 1 / (1 + al)
        return np.array([[al * inv, 0.0, -al * inv, 1.0, -2 * np.cos(w0) * inv, (1 - al) * inv]])

    acc = np.zeros_like(g["x"])
    for sos in (bpf(500, 1.414), bpf(2000, 1.414), lp):
        acc += oracle.sos_cascade(g["x"], sos)[0]
    assert rel_to_max(acc, g["comb"]) < 3e-7


def test_fir_matches_reference_golden():
    g = golden("fir.npz")
    y = oracle.fir_causal(g["x"], g["taps"])
    assert rel_to_max(y, g["y_fft"]) < 5e-6  # reference evaluates in float32 FFTs
    assert rel_to_max(y, g["y_direct"]) < 5e-6
    assert rel_to_max(oracle.fir_causal(g["x"], g["des_b"].astype(np.float32)), g["y_des"]) < 5e-6
    # the block-structured restatement of fft_conv1d
    yc = oracle.fft_conv1d(g["xc"], g["kern"])
    assert yc.shape == g["yc"].shape and rel_to_max(yc, g["yc"]) < 5e-6
    ycp = oracle.fft_conv1d(g["xc"], g["kern"], padding=(5, 10))
    assert ycp.shape == g["yc_pad"].shape and rel_to_max(ycp, g["yc_pad"]) < 5e-6
    with pytest.raises(RuntimeError, match="at least as large"):
        oracle.fft_conv1d(g["xc"][..., :10], g["kern"])
    with pytest.raises(RuntimeError, match="Block ratio"):
        oracle.fft_conv1d(g["xc"], g["kern"], block_ratio=0.5)


def test_delay_matches_reference_golden():
    g = golden("delay.npz")
    assert rel_to_max(oracle.delay_line(g["x"], 100, 0.5, 0.8), g["y"]) < 2e-7
    np.testing.assert_allclose(oracle.delay_line(g["x64"], 333, 0.7, 0.25), g["y64"], rtol=1e-14, atol=1e-15)
    short = g["x"][:, :50]
    assert np.array_equal(oracle.delay_line(short, 100, 0.5, 0.5), short)


def test_oracle_vs_live_reference_when_available():
    """In the build container the unmodified reference extension is loadable: compare live."""
    from oracle import ref_loader

    if not ref_loader.have_ref_ext():
        pytest.skip("oracle/_ref/torchfx_ext.so not built")
    import torch

    ext = ref_loader.load_ref_ext()
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 4000))
    sos = sps.ellip(6, 0.5, 50, 0.25, output="sos")
    K = sos.shape[0]
    sx = rng.standard_normal((K, 3, 2))
    sy = rng.standard_normal((K, 3, 2))
    yr, sxr, syr = ext.sos_forward(torch.from_numpy(x), torch.from_numpy(sos), torch.from_numpy(sos), torch.from_numpy(sx), torch.from_numpy(sy))
    y, sx1, sy1 = oracle.sos_cascade(x, sos, sx, sy)
    np.testing.assert_allclose(y, yr.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(sx1, sxr.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(sy1, syr.numpy(), rtol=1e-10, atol=1e-12)
