"""``fft_conv1d``: the reference's overlap-save convolution entry point, served natively.

Reference: src/torchfx/filter/_fftconv.py:70-141 -- cross-correlation of ``x`` ``[B, C, T]``
(after ``F.pad(x, padding)``) with ``kernel`` ``[..., K]``, output length ``L - K + 1``,
computed there with torch.fft in blocks of ``int(K * block_ratio)``.  The block size only
shapes the reference's intermediate tensors, not the result, so it is validated and
otherwise ignored; the sum itself is evaluated by the library's FIR kernels.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .fir import fir_causal


def pad_to(tensor: Tensor, target_length: int) -> Tensor:
    """Zero-extend the last axis to ``target_length`` samples (reference _fftconv.py:23-29)."""
    extra = target_length - tensor.shape[-1]
    if extra == 0:
        return tensor
    if extra < 0:  # F.pad semantics: a negative amount crops
        return tensor[..., :target_length]
    out = tensor.new_zeros(tensor.shape[:-1] + (target_length,))
    out[..., : tensor.shape[-1]] = tensor
    return out


def unfold(x: Tensor, kernel_size: int, stride: int) -> Tensor:
    """Overlapping frames of the last axis as a zero-copy view ``[*, F, kernel_size]``,
    ``F = 1 + ceil((max(T, kernel_size) - kernel_size) / stride)``; the tail is zero-padded
    so the last frame is complete (reference _fftconv.py:32-67).  The native overlap-save
    kernels never materialise frames; this helper exists for callers of the reference API."""
    length = x.shape[-1]
    n_frames = -(-(max(length, kernel_size) - kernel_size) // stride) + 1
    padded = pad_to(x, (n_frames - 1) * stride + kernel_size).contiguous()
    return padded.unfold(-1, kernel_size, stride)


def fft_conv1d(x: Tensor, kernel: Tensor, padding: tuple[int, int] = (0, 0), block_ratio: float = 5.0) -> Tensor:
    if x.ndim != 3:
        raise ValueError(f"expected [B, C, T], got {tuple(x.shape)}")
    batch, channels, length = x.shape
    length += padding[0] + padding[1]
    k = kernel.shape[-1]
    if length < k:
        raise RuntimeError(
            f"Input should be at least as large as the kernel size {k}, but it is only {length} samples long."
        )
    if block_ratio < 1:
        raise RuntimeError("Block ratio must be greater than 1.")
    # out[n] = sum_i w[i] * xp[n + i]  ==  causal FIR with taps w reversed, delayed by K-1
    xp = torch.nn.functional.pad(x, padding) if (padding[0] or padding[1]) else x
    taps = kernel.reshape(-1, k)[0].flip(0)
    y = fir_causal(xp.reshape(batch * channels, length), taps)
    return y[:, k - 1 :].reshape(batch, channels, length - k + 1)
