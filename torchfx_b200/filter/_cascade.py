"""Shared forward path of every SOS filter (IIR, Biquad, FusedSOSCascade).

Reference: ``_sos_cascade_forward`` (src/torchfx/filter/iir.py:84-184): accept ``[T]``,
``[C, T]`` or ``[B, C, T]`` (B*C flattened, :119-126); keep per-section DF1 state
``[K, C, 2]`` float64 across calls and re-allocate it when the channel count changes
(:135-144); return the input's dtype (:165,176).  The reference's K == 1 CUDA special case
(:149-172) is not needed: one kernel serves every K.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import _ops


def run_sos_cascade(
    x: Tensor, sos_cpu: Tensor, state_x: Tensor | None, state_y: Tensor | None
) -> tuple[Tensor, Tensor, Tensor]:
    """Returns ``(y, state_x, state_y)``; the state tensors are owned by the caller module
    and updated in place by the native call."""
    shape = x.shape
    if x.ndim == 1:
        x2 = x.unsqueeze(0)
    elif x.ndim == 2:
        x2 = x
    elif x.ndim == 3:
        x2 = x.reshape(shape[0] * shape[1], shape[2])
    else:
        raise ValueError("Input must be of shape [T], [C, T], or [B, C, T]")
    C = x2.shape[0]
    K = sos_cpu.shape[0]
    if state_x is None or state_y is None or state_x.shape[1] != C or state_x.shape[0] != K:
        state_x = torch.zeros(K, C, 2, device=x2.device, dtype=torch.float64)
        state_y = torch.zeros(K, C, 2, device=x2.device, dtype=torch.float64)
    elif state_x.device != x2.device:
        state_x = state_x.to(x2.device)
        state_y = state_y.to(x2.device)
    y = _ops.sos_cascade_(x2, sos_cpu, state_x, state_y)
    return y.reshape(shape), state_x, state_y
