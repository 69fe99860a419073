"""``FX`` base class and the effects that touch the native layer.

Reference: src/torchfx/effect.py:139-258 (``FX``: an abstract ``nn.Module`` whose ``|``
builds a ``FilterChain``), :261-383 (``Gain``) and :845-931 (``Reverb`` ->
``_ops.delay_line_forward``).  Only what the filter path needs is rebuilt (SURVEY.md 2,
rows 9-10); Normalize / Delay strategies are out of scope.
"""
from __future__ import annotations

import abc
import math

import torch
from torch import Tensor, nn


class FX(nn.Module, abc.ABC):
    """Abstract effect: a module mapping a waveform tensor to a waveform tensor."""

    @abc.abstractmethod
    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)

    @abc.abstractmethod
    def forward(self, x: Tensor) -> Tensor: ...

    def __or__(self, other: nn.Module) -> nn.Sequential:
        # f1 | f2 -> FilterChain, flattened (reference effect.py:253-258, chain.py:52-59)
        if not isinstance(other, nn.Module):
            return NotImplemented
        from .chain import FilterChain

        return FilterChain(self, other)


class Gain(FX):
    """Amplitude / dB / power gain (reference effect.py:261-383).  In the reference a Gain always
    breaks an IIR run in ``Wave._materialize`` (tests/test_chain_fusion.py:102-120); here a
    non-clamping Gain next to a fused run of >= 2 filters is folded into that run's
    coefficients (``torchfx_b200.wave.FOLD_GAIN``), which leaves the results and the set of
    statefully-run modules unchanged; next to a lone filter it stays a step of its own."""

    def __init__(self, gain: float, gain_type: str = "amplitude", clamp: bool = False) -> None:
        super().__init__()
        if gain_type in ("amplitude", "power") and gain < 0:
            raise ValueError("If gain_type = amplitude or power, gain must be positive.")
        self.gain = gain
        self.gain_type = gain_type
        self.clamp = clamp

    def linear_gain(self) -> float:
        if self.gain_type == "amplitude":
            return float(self.gain)
        if self.gain_type == "db":
            return 10.0 ** (self.gain / 20.0)
        if self.gain_type == "power":
            return 10.0 ** (10.0 * math.log10(self.gain) / 20.0)
        return 1.0

    @torch.no_grad()
    def forward(self, waveform: Tensor) -> Tensor:
        g = self.linear_gain()
        if g != 1.0:
            waveform = waveform * g
        if self.clamp:
            waveform = torch.clamp(waveform, -1.0, 1.0)
        return waveform


class Reverb(FX):
    """Single feed-forward echo, y[n] = x[n] + mix*decay*x[n-delay], through the native
    delay-line op (reference effect.py:845-931; delay_cpu.cpp:17-85)."""

    def __init__(self, delay: int = 4410, decay: float = 0.5, mix: float = 0.5) -> None:
        super().__init__()
        assert delay > 0, "Delay must be positive."
        assert 0 < decay < 1, "Decay must be between 0 and 1."
        assert 0 <= mix <= 1, "Mix must be between 0 and 1."
        self.delay = delay
        self.decay = decay
        self.mix = mix

    @torch.no_grad()
    def forward(self, waveform: Tensor) -> Tensor:
        if waveform.size(-1) <= self.delay:
            return waveform
        from ._ops import delay_line_forward

        if waveform.ndim <= 2:
            return delay_line_forward(waveform, self.delay, self.decay, self.mix)
        flat = waveform.reshape(-1, waveform.shape[-1])
        return delay_line_forward(flat, self.delay, self.decay, self.mix).reshape(waveform.shape)
