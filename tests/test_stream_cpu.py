"""Host-side tests of the chunked driver and the WAV codec.

Reference behaviour: src/torchfx/realtime/stream.py (constructor validation :70-83,
Nyquist check :146-156, overlap trimming :240-246) and src/torchfx/wave.py:406-576
(from_file metadata, save subtype mapping); reference tests tests/test_realtime_stream.py.
"""
import struct
import wave as std_wave

import numpy as np
import pytest
import scipy.signal as sps
import torch

import torchfx_b200 as fx
from torchfx_b200 import _wavio
from torchfx_b200.filter import HiButterworth, LoButterworth
from torchfx_b200.realtime import StreamProcessor

FS = 48000


def _signal(c=2, t=20000, seed=0):
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(c, t, generator=g)


@pytest.mark.parametrize("subtype,tol", [("FLOAT", 0.0), ("DOUBLE", 0.0), ("PCM_16", 1.0 / 32767), ("PCM_24", 1.0 / 8388607),
                                         ("PCM_32", 1e-7), ("PCM_U8", 1.0 / 127)])
def test_wav_roundtrip(tmp_path, subtype, tol):
    x = _signal(3, 5000).numpy().T
    p = tmp_path / "a.wav"
    _wavio.write(p, x, FS, subtype)
    meta = _wavio.info(p)
    assert (meta.samplerate, meta.frames, meta.channels, meta.subtype) == (FS, 5000, 3, subtype)
    y, fs = _wavio.read(p)
    assert fs == FS and y.shape == x.shape and y.dtype == np.float32
    assert np.abs(y - x).max() <= tol
    part, _ = _wavio.read(p, 1000, 1500)
    np.testing.assert_array_equal(part, y[1000:1500])
    tail, _ = _wavio.read(p, 4900, 9999)
    assert tail.shape == (100, 3)


def test_wav_matches_stdlib_pcm16(tmp_path):
    """Known-answer check against Python's own WAV codec (16-bit PCM)."""
    x = (np.arange(-2000, 2000, dtype=np.int16).reshape(-1, 2) * 7).astype(np.int16)
    p = tmp_path / "std.wav"
    with std_wave.open(str(p), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(x.tobytes())
    y, fs = _wavio.read(p)
    assert fs == 44100
    np.testing.assert_array_equal(y, x.astype(np.float32) / 32768.0)
    q = tmp_path / "ours.wav"
    _wavio.write(q, y, 44100, "PCM_16")
    with std_wave.open(str(q), "rb") as r:
        assert (r.getnchannels(), r.getsampwidth(), r.getframerate(), r.getnframes()) == (2, 2, 44100, 2000)
        back = np.frombuffer(r.readframes(2000), dtype=np.int16).reshape(-1, 2)
    # float -> PCM scales by 32767 (libsndfile): within one LSB of the original
    assert np.abs(back.astype(np.int32) - x.astype(np.int32)).max() <= 1


def test_wav_extensible_header_and_errors(tmp_path):
    x = _signal(2, 100).numpy().T.astype("<f4")
    body = x.tobytes()
    fmt = struct.pack("<HHIIHH", 0xFFFE, 2, FS, FS * 8, 8, 32) + struct.pack("<HHI", 22, 32, 3) + struct.pack(
        "<H", 3) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    junk = b"LIST" + struct.pack("<I", 3) + b"abc\0"
    riff = b"WAVE" + junk + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(body)) + body
    p = tmp_path / "ext.wav"
    p.write_bytes(b"RIFF" + struct.pack("<I", len(riff)) + riff)
    y, _ = _wavio.read(p)
    np.testing.assert_array_equal(y, x)
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a wave file at all")
    with pytest.raises(ValueError, match="not a RIFF/WAVE"):
        _wavio.info(bad)
    with pytest.raises(ValueError, match="unsupported WAV subtype"):
        _wavio.WavWriter(tmp_path / "x.wav", FS, 1, "VORBIS")


def test_wave_from_file_and_save(tmp_path):
    x = _signal(2, 3000)
    w = fx.Wave(x, FS)
    p = tmp_path / "sub" / "out.wav"
    w.save(p, encoding="PCM_F", bits_per_sample=32)
    r = fx.Wave.from_file(p)
    assert r.fs == FS and r.metadata == {"num_frames": 3000, "num_channels": 2, "subtype": "FLOAT", "format": "WAV"}
    torch.testing.assert_close(r.ys, x, atol=0, rtol=0)
    part = fx.Wave.from_file(p, frame_offset=100, num_frames=50)
    torch.testing.assert_close(part.ys, x[:, 100:150], atol=0, rtol=0)
    w.save(tmp_path / "pcm.wav")  # libsndfile's WAV default: PCM_16
    assert fx.Wave.from_file(tmp_path / "pcm.wav").metadata["subtype"] == "PCM_16"
    w.save(tmp_path / "p24.wav", bits_per_sample=24)
    r24 = fx.Wave.from_file(tmp_path / "p24.wav")
    assert r24.metadata["subtype"] == "PCM_24" and (r24.ys - x).abs().max() < 2e-7
    # the pipe works on a loaded wave and saves lazily-materialised data
    (fx.Wave.from_file(p) | LoButterworth(4000, order=2)).save(tmp_path / "f.wav", encoding="PCM_F", bits_per_sample=32)
    ref = sps.sosfilt(sps.butter(2, 4000, fs=FS, output="sos"), x.double().numpy(), axis=-1)
    np.testing.assert_allclose(fx.Wave.from_file(tmp_path / "f.wav").ys.numpy(), ref, atol=2e-6)


def test_stream_processor_validation():
    with pytest.raises(ValueError, match="chunk_size must be positive"):
        StreamProcessor([fx.Gain(0.5)], chunk_size=0)
    with pytest.raises(ValueError, match="non-negative"):
        StreamProcessor([fx.Gain(0.5)], overlap=-1)
    with pytest.raises(ValueError, match="must be less than chunk_size"):
        StreamProcessor([fx.Gain(0.5)], chunk_size=16, overlap=16)
    with pytest.raises(TypeError, match="must inherit from FX"):
        StreamProcessor([torch.nn.Identity()])
    p = StreamProcessor(fx.FilterChain(fx.Gain(0.5), fx.Gain(2.0)), chunk_size=128, overlap=8)
    assert p.chunk_size == 128 and p.overlap == 8 and len(p.effects) == 2


def test_stream_chunks_equal_one_shot_cpu(tmp_path):
    """Stateful filters carried across chunks == one unbroken call (reference contract
    filter/iir.py:135-144), through files and through host tensors."""
    x = _signal(2, 30000, seed=5)
    src = tmp_path / "in.wav"
    fx.Wave(x, FS).save(src, encoding="PCM_F", bits_per_sample=32)
    ref = sps.sosfilt(sps.butter(3, 200, "high", fs=FS, output="sos"),
                      0.5 * sps.sosfilt(sps.butter(4, 3000, fs=FS, output="sos"), x.double().numpy(), axis=-1), axis=-1)
    proc = StreamProcessor([LoButterworth(3000, order=4), fx.Gain(0.5), HiButterworth(200, order=3)], chunk_size=4096)
    chunks = list(proc.process_chunks(src))
    assert [c.shape[1] for c in chunks] == [4096] * 7 + [30000 - 7 * 4096]
    np.testing.assert_allclose(torch.cat(chunks, dim=1).numpy(), ref, atol=2e-6)
    proc.reset_state()
    proc.process_file(src, tmp_path / "o" / "out.wav")
    out = fx.Wave.from_file(tmp_path / "o" / "out.wav")
    assert out.metadata["subtype"] == "FLOAT" and out.fs == FS
    np.testing.assert_allclose(out.ys.numpy(), ref, atol=2e-6)
    proc.reset_state()
    np.testing.assert_allclose(proc.process_tensor(x, fs=FS).numpy(), ref, atol=2e-6)
    # on "cpu" the fused streaming driver is never selected
    assert proc._fused_cascade() is None


def test_stream_overlap_trims_and_nyquist(tmp_path):
    x = _signal(1, 1000, seed=6)
    src = tmp_path / "in.wav"
    fx.Wave(x, 8000).save(src, encoding="PCM_F", bits_per_sample=32)
    proc = StreamProcessor([fx.Gain(2.0)], chunk_size=256, overlap=64)
    chunks = list(proc.process_chunks(src))
    # hop = 192: first chunk whole, later chunks lose their first `overlap` samples (stream.py:240-246)
    assert chunks[0].shape[1] == 256 and chunks[1].shape[1] == 192
    np.testing.assert_allclose(chunks[1].numpy(), 2.0 * x[:, 192 + 64: 192 + 256].numpy(), atol=1e-7)
    with pytest.raises(ValueError, match="below the Nyquist"):
        list(StreamProcessor([LoButterworth(5000, order=2)]).process_chunks(src))


def test_fused_stream_selection_rules():
    lo, hi = LoButterworth(3000, order=2, fs=FS), HiButterworth(100, order=2, fs=FS)
    p = StreamProcessor([lo, fx.Gain(0.5), hi], device="cuda")
    fused = p._fused_cascade()
    assert fused is not None and fused._num_sections == 2 and fused.gain == 0.5
    assert p._fused_cascade() is fused  # cached
    assert StreamProcessor([lo, hi], device="cuda", chunk_size=64, overlap=8)._fused_cascade() is None
    assert StreamProcessor([lo, fx.Gain(2.0, clamp=True)], device="cuda")._fused_cascade() is None
    assert StreamProcessor([fx.Gain(2.0)], device="cuda")._fused_cascade() is None
    assert StreamProcessor([lo, fx.Reverb(100)], device="cuda")._fused_cascade() is None
