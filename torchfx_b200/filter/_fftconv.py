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
