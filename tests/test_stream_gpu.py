"""GPU parity of the host-buffer streaming driver behind StreamProcessor."""
import numpy as np
import pytest
import scipy.signal as sps
import torch

import torchfx_b200 as fx
from oracle import oracle
from torchfx_b200 import _ops
from torchfx_b200.filter import HiButterworth, LoButterworth, ParametricEQ
from torchfx_b200.realtime import StreamProcessor

from conftest import rel_to_max

pytestmark = pytest.mark.gpu
FS = 48000


def _signal(c, t, seed=0):
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(c, t, generator=g)


@pytest.mark.parametrize("pinned", [False, True])
def test_fused_stream_matches_oracle_and_carries_state(pinned):
    x = _signal(6, 300_000, seed=1)
    if pinned:
        x = x.pin_memory()
    effects = [LoButterworth(3000, order=4), fx.Gain(0.5), HiButterworth(200, order=3), ParametricEQ(1000, q=2.0, gain=3.0)]
    proc = StreamProcessor(effects, device="cuda")
    y = proc.process_tensor(x, fs=FS)
    fused = proc._fused_cascade()
    assert fused is not None and y.shape == x.shape and not y.is_cuda and y.is_pinned() == pinned
    # independent oracle for the folded gain: cascade -> x g -> cascade, each filter's OWN coefficients
    mid, _, _ = oracle.sos_cascade(x.numpy(), effects[0]._sos.numpy())
    mid = (mid.astype(np.float64) * 0.5)
    ref, _, _ = oracle.sos_cascade(mid, np.vstack([effects[2]._sos.numpy(), effects[3]._sos.numpy()]))
    assert fused.gain == 0.5 and fused._num_sections == 2 + 2 + 1
    assert rel_to_max(y.numpy(), ref) < 1e-5
    # two halves with carried state == one call
    proc.reset_state()
    a = proc.process_tensor(x[:, :123_457])
    b = proc.process_tensor(x[:, 123_457:])
    assert rel_to_max(torch.cat([a, b], dim=1).numpy(), ref) < 1e-5
    # the reference's generic per-chunk loop (modules own their state) gives the same signal
    generic = StreamProcessor([LoButterworth(3000, order=4), fx.Gain(0.5), HiButterworth(200, order=3),
                               ParametricEQ(1000, q=2.0, gain=3.0)], chunk_size=65536, overlap=0, device="cuda")
    generic._fused_cascade = lambda: None
    z = generic.process_tensor(x, fs=FS)
    assert rel_to_max(z.numpy(), ref) < 1e-5


def test_host_driver_small_chunks_and_state_roundtrip():
    x = _signal(3, 70_001, seed=2)
    sos = torch.as_tensor(sps.butter(6, 4000, fs=FS, output="sos"))
    K = sos.shape[0]
    sx = torch.zeros(K, 3, 2, dtype=torch.float64)
    sy = torch.zeros(K, 3, 2, dtype=torch.float64)
    y = _ops.sos_cascade_host_(x, sos, sx, sy, chunk=16384)  # 5 device chunks, ragged tail
    ref, rsx, rsy = oracle.sos_cascade(x.numpy(), sos.numpy())
    assert rel_to_max(y.numpy(), ref) < 1e-5
    np.testing.assert_allclose(sx.numpy(), rsx, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sy.numpy(), rsy, rtol=1e-4, atol=1e-6)
    with pytest.raises(ValueError, match="HOST tensor"):
        _ops.sos_cascade_host_(x.cuda(), sos)


def test_process_file_on_gpu(tmp_path):
    x = _signal(2, 200_000, seed=3)
    src = tmp_path / "in.wav"
    fx.Wave(x, FS).save(src, encoding="PCM_F", bits_per_sample=32)
    proc = StreamProcessor([LoButterworth(5000, order=8)], device="cuda")
    proc.process_file(src, tmp_path / "out.wav")
    out = fx.Wave.from_file(tmp_path / "out.wav")
    ref, _, _ = oracle.sos_cascade(x.numpy(), proc._fused_cascade()._sos.numpy())
    assert rel_to_max(out.ys.numpy(), ref) < 1e-5
