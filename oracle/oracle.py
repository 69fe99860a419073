"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/oracle.c + numpy).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  torchfx_b200/ never does.

Every function restates a piece of the reference (matteospanio/torchfx @ 27e65b0) and
cites it; parity of the oracle itself is pinned by tests/test_oracle.py against golden
vectors generated from the unmodified reference (oracle/make_golden.py) and against
scipy.signal (the oracle the reference's own tests use).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from ctypes import c_double, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build() -> str:
    """Compile oracle.c with gcc (no CUDA, no torch)."""
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        P = c_void_p
        lib.oracle_biquad_df1_f64.argtypes = [P, P, c_int64, c_int64, c_double, c_double, c_double, c_double, c_double, P, P]
        lib.oracle_sos_df1_f64.argtypes = [P, P, c_int64, c_int64, P, c_int64, P, P]
        lib.oracle_sos_df1_f32io.argtypes = [P, P, c_int64, c_int64, c_int64, c_int64, P, c_int64, P, P]
        lib.oracle_delay_line_f32.argtypes = [P, P, c_int64, c_int64, c_int64, c_double, c_double]
        lib.oracle_delay_line_f64.argtypes = [P, P, c_int64, c_int64, c_int64, c_double, c_double]
        lib.oracle_fir_causal_f32.argtypes = [P, P, c_int64, c_int64, P, c_int64]
        for name in ("oracle_biquad_df1_f64", "oracle_sos_df1_f64", "oracle_sos_df1_f32io", "oracle_delay_line_f32",
                     "oracle_delay_line_f64", "oracle_fir_causal_f32"):
            getattr(lib, name).restype = None
        _lib = lib
    return _lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def _state(s, shape):
    if s is None:
        return np.zeros(shape, dtype=np.float64)
    s = np.array(s, dtype=np.float64, order="C", copy=True)
    assert s.shape == tuple(shape), (s.shape, shape)
    return s


def sos_cascade(x: np.ndarray, sos: np.ndarray, state_x=None, state_y=None):
    """Reference contract of one SOS-cascade call on ``x`` [C, T] (float32 or float64):
    f64 DF1 cascade (cpu/iir_cpu.cpp:64-159), result cast to the input dtype
    (filter/iir.py:176).  Returns ``(y, new_state_x, new_state_y)``, states [K, C, 2]."""
    lib = _load()
    x = np.ascontiguousarray(x)
    assert x.ndim == 2 and x.dtype in (np.float32, np.float64)
    sos = np.ascontiguousarray(sos, dtype=np.float64)
    C, T = x.shape
    K = sos.shape[0]
    sx = _state(state_x, (K, C, 2))
    sy = _state(state_y, (K, C, 2))
    y = np.empty_like(x)
    if x.dtype == np.float32:
        lib.oracle_sos_df1_f32io(_p(x), _p(y), C, T, T, T, _p(sos), K, _p(sx), _p(sy))
    else:
        lib.oracle_sos_df1_f64(_p(x), _p(y), C, T, _p(sos), K, _p(sx), _p(sy))
    return y, sx, sy


def biquad(x: np.ndarray, b, a1: float, a2: float, state_x=None, state_y=None):
    """cpu/iir_cpu.cpp:10-62 on float64 ``x`` [C, T]; states [C, 2]."""
    lib = _load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    C, T = x.shape
    sx = _state(state_x, (C, 2))
    sy = _state(state_y, (C, 2))
    y = np.empty_like(x)
    lib.oracle_biquad_df1_f64(_p(x), _p(y), C, T, float(b[0]), float(b[1]), float(b[2]), float(a1), float(a2), _p(sx), _p(sy))
    return y, sx, sy


def delay_line(x: np.ndarray, delay: int, decay: float, mix: float) -> np.ndarray:
    """cpu/delay_cpu.cpp:17-85."""
    lib = _load()
    x = np.ascontiguousarray(x)
    x2 = x[None, :] if x.ndim == 1 else x
    y = np.empty_like(x2)
    fn = lib.oracle_delay_line_f32 if x.dtype == np.float32 else lib.oracle_delay_line_f64
    fn(_p(x2), _p(y), x2.shape[0], x2.shape[1], int(delay), float(decay), float(mix))
    return y.reshape(x.shape)


def fir_causal(x: np.ndarray, b: np.ndarray) -> np.ndarray:
    """y[n] = sum_j b[j] x[n-j], f64 accumulate, f32 in/out (what filter/fir.py:526-579 computes)."""
    lib = _load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    y = np.empty_like(x)
    lib.oracle_fir_causal_f32(_p(x), _p(y), x.shape[0], x.shape[1], _p(b), b.shape[0])
    return y


def fft_conv1d(x: np.ndarray, kernel: np.ndarray, padding=(0, 0), block_ratio: float = 5.0) -> np.ndarray:
    """numpy restatement of the reference's overlap-save cross-correlation,
    filter/_fftconv.py:107-141: block = min(int(K*block_ratio), L), hop = block-K+1,
    frames (unfold, :32-67) -> rfft * conj(rfft(kernel)) -> irfft, keep [:hop] of each
    frame, trim to L-K+1.  x is [B, C, T]; computed in the dtype of x like torch.fft."""
    x = np.pad(x, [(0, 0), (0, 0), tuple(padding)])
    B, C, L = x.shape
    K = kernel.shape[-1]
    if L < K:
        raise RuntimeError(f"Input should be at least as large as the kernel size {K}, but it is only {L} samples long.")
    if block_ratio < 1:
        raise RuntimeError("Block ratio must be greater than 1.")
    block = min(int(K * block_ratio), L)
    hop = block - K + 1
    kz = np.fft.rfft(np.pad(kernel.reshape(-1)[:K], (0, block - K)).astype(x.dtype))
    n_frames = math.ceil((max(L, block) - block) / hop) + 1
    tgt = (n_frames - 1) * hop + block
    xp = np.pad(x, [(0, 0), (0, 0), (0, tgt - L)])
    idx = np.arange(n_frames)[:, None] * hop + np.arange(block)[None, :]
    frames = xp[..., idx]  # [B, C, F, block]
    out = np.fft.irfft(np.fft.rfft(frames, axis=-1) * np.conj(kz), n=block, axis=-1)[..., :hop]
    out = out.reshape(B, C, -1)[..., : L - K + 1]
    return out.astype(x.dtype)


def filterbank_stack(x: np.ndarray, sos_bands: np.ndarray) -> np.ndarray:
    """LogFilterBank semantics (filter/filterbank.py:183-185): band b is an independent
    cascade over the same x; outputs stacked [N, C, T]."""
    return np.stack([sos_cascade(x, sos_bands[b])[0] for b in range(sos_bands.shape[0])], axis=0)


def filterbank_sum(x: np.ndarray, sos_bands: np.ndarray) -> np.ndarray:
    """ParallelFilterCombination semantics (filter/__base.py:1019-1026): zeros_like(x) then
    += each child's output, in child order, in the dtype of x."""
    acc = np.zeros_like(x)
    for b in range(sos_bands.shape[0]):
        acc += sos_cascade(x, sos_bands[b])[0]
    return acc
