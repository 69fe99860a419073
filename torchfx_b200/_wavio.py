"""RIFF/WAVE reader and writer in numpy.

The reference does all file I/O through ``soundfile`` (src/torchfx/wave.py:406-470 read,
:472-576 write; src/torchfx/realtime/stream.py:160-255 chunked), which is not installed
in this image.  WAV is the format the streaming path actually moves, so it is handled
here directly: PCM 8/16/24/32-bit, IEEE float 32/64, plain and WAVE_FORMAT_EXTENSIBLE
headers, partial reads by frame range (what ``sf.read(start=, stop=)`` gives the chunked
driver).  Sample scaling follows libsndfile's float conventions: PCM -> float divides by
2^(bits-1); float -> PCM multiplies by 2^(bits-1) - 1 and rounds to nearest.
Other containers (FLAC, OGG) are delegated to ``soundfile`` when it is importable.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

_PCM, _FLOAT, _EXTENSIBLE = 1, 3, 0xFFFE


@dataclass
class WavInfo:
    samplerate: int
    frames: int
    channels: int
    subtype: str
    format: str = "WAV"
    # private: where the samples live
    data_offset: int = 0
    sample_bytes: int = 0
    is_float: bool = False


def _subtype(is_float: bool, bits: int) -> str:
    if is_float:
        return "FLOAT" if bits == 32 else "DOUBLE"
    return "PCM_U8" if bits == 8 else f"PCM_{bits}"


def info(path: str | os.PathLike) -> WavInfo:
    with open(path, "rb") as fh:
        head = fh.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        while True:
            hdr = fh.read(8)
            if len(hdr) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if cid == b"fmt ":
                body = fh.read(size + (size & 1))
                tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == _EXTENSIBLE and size >= 26:
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, ch, rate, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                tag, ch, rate, bits = fmt
                if tag not in (_PCM, _FLOAT) or bits not in (8, 16, 24, 32, 64) or ch < 1:
                    raise ValueError(f"{path}: unsupported WAV encoding (tag {tag}, {bits} bit)")
                offset = fh.tell()
                remaining = os.fstat(fh.fileno()).st_size - offset
                if size == 0xFFFFFFFF or size > remaining:  # streamed / truncated header
                    size = remaining
                sb = bits // 8
                return WavInfo(rate, size // (sb * ch), ch, _subtype(tag == _FLOAT, bits), "WAV", offset, sb,
                               tag == _FLOAT)
            else:
                fh.seek(size + (size & 1), os.SEEK_CUR)


def _decode(raw: bytes, meta: WavInfo) -> np.ndarray:
    sb = meta.sample_bytes
    if meta.is_float:
        a = np.frombuffer(raw, dtype="<f4" if sb == 4 else "<f8").astype(np.float32)
    elif sb == 1:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif sb == 2:
        a = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif sb == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = (v ^ 0x800000) - 0x800000  # sign-extend 24 -> 32
        a = v.astype(np.float32) / 8388608.0
    else:
        a = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    return a.reshape(-1, meta.channels)


def read(path: str | os.PathLike, start: int = 0, stop: int | None = None,
         meta: WavInfo | None = None) -> tuple[np.ndarray, int]:
    """Frames ``[start, stop)`` as float32 ``[frames, channels]`` plus the sample rate."""
    meta = meta or info(path)
    stop = meta.frames if stop is None else min(stop, meta.frames)
    start = max(0, min(start, stop))
    fb = meta.sample_bytes * meta.channels
    with open(path, "rb") as fh:
        fh.seek(meta.data_offset + start * fb)
        raw = fh.read((stop - start) * fb)
    return _decode(raw, meta), meta.samplerate


def _encode(frames: np.ndarray, subtype: str) -> bytes:
    a = np.ascontiguousarray(frames)
    if subtype == "FLOAT":
        return a.astype("<f4").tobytes()
    if subtype == "DOUBLE":
        return a.astype("<f8").tobytes()
    a = a.astype(np.float64)
    if subtype == "PCM_U8":
        return (np.clip(np.rint(a * 127.0), -128, 127) + 128).astype(np.uint8).tobytes()
    bits = int(subtype.split("_")[1])
    full = float((1 << (bits - 1)) - 1)
    v = np.clip(np.rint(a * full), -full - 1, full).astype(np.int64)
    if bits == 16:
        return v.astype("<i2").tobytes()
    if bits == 32:
        return v.astype("<i4").tobytes()
    if bits == 24:
        u = (v & 0xFFFFFF).astype(np.uint32).reshape(-1)
        out = np.empty((u.size, 3), dtype=np.uint8)
        out[:, 0] = u & 0xFF
        out[:, 1] = (u >> 8) & 0xFF
        out[:, 2] = (u >> 16) & 0xFF
        return out.tobytes()
    raise ValueError(f"unsupported WAV subtype {subtype!r}")


_SUBTYPES = {"PCM_U8": (8, False), "PCM_16": (16, False), "PCM_24": (24, False), "PCM_32": (32, False),
             "FLOAT": (32, True), "DOUBLE": (64, True)}


class WavWriter:
    """Progressive writer (what ``sf.SoundFile(mode="w")`` is to the reference's chunked
    driver, realtime/stream.py:196-205): ``write`` appends ``[frames, channels]`` blocks, the
    RIFF sizes are patched on close."""

    def __init__(self, path: str | os.PathLike, samplerate: int, channels: int, subtype: str | None = None) -> None:
        subtype = subtype or "PCM_16"  # libsndfile's default for WAV
        if subtype not in _SUBTYPES:
            raise ValueError(f"unsupported WAV subtype {subtype!r}")
        self.subtype = subtype
        self.channels = channels
        bits, is_float = _SUBTYPES[subtype]
        self._fh = open(path, "wb")
        sb = bits // 8
        fmt = struct.pack("<HHIIHH", _FLOAT if is_float else _PCM, channels, samplerate, samplerate * channels * sb,
                          channels * sb, bits)
        self._fh.write(b"RIFF\0\0\0\0WAVEfmt " + struct.pack("<I", len(fmt)) + fmt + b"data\0\0\0\0")
        self._data_start = self._fh.tell()

    def write(self, frames: np.ndarray) -> None:
        frames = np.asarray(frames)
        if frames.ndim == 1:
            frames = frames[:, None]
        if frames.shape[1] != self.channels:
            raise ValueError(f"expected {self.channels} channels, got {frames.shape[1]}")
        self._fh.write(_encode(frames, self.subtype))

    def close(self) -> None:
        if self._fh.closed:
            return
        end = self._fh.tell()
        nbytes = end - self._data_start
        if nbytes & 1:
            self._fh.write(b"\0")
            end += 1
        self._fh.seek(4)
        self._fh.write(struct.pack("<I", min(end - 8, 0xFFFFFFFF)))
        self._fh.seek(self._data_start - 4)
        self._fh.write(struct.pack("<I", min(nbytes, 0xFFFFFFFF)))
        self._fh.close()

    def __enter__(self) -> "WavWriter":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


def write(path: str | os.PathLike, frames: np.ndarray, samplerate: int, subtype: str | None = None) -> None:
    frames = np.asarray(frames)
    if frames.ndim == 1:
        frames = frames[:, None]
    with WavWriter(path, samplerate, frames.shape[1], subtype) as w:
        w.write(frames)


def is_wav_path(path: str | os.PathLike, format: str | None = None) -> bool:  # noqa: A002
    if format is not None:
        return format.upper() == "WAV"
    return os.path.splitext(str(path))[1].lower() in (".wav", ".wave", "")
