"""Literal types used by the filter signatures (reference: src/torchfx/typing.py)."""
from __future__ import annotations

import typing as tp

import torch

Device = tp.Union[str, torch.device]
FilterOrderScale = tp.Literal["db", "linear"]
WindowType = tp.Literal[
    "hann", "hamming", "blackman", "kaiser", "boxcar", "bartlett", "flattop", "parzen", "bohman", "nuttall", "barthann"
]
Second = float
Millisecond = float
