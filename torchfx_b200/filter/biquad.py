"""Second-order (biquad) filters with closed-form coefficients.

Reference: src/torchfx/filter/biquad.py -- ``Biquad`` base (:73-236: ``[1, 6]`` SOS row,
state attributes, ``forward`` -> the shared SOS cascade, ``reset_state`` keeps the
coefficients :198-206) and the six AudioNoise/RBJ designs LPF/HPF/Notch/BPF/BPFPeak/AllPass
(:269-509), pinned to 1e-10 by the reference's tests/test_biquad.py:42-125.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from ._base import AbstractFilter
from ._cascade import run_sos_cascade


class Biquad(AbstractFilter):
    """Base class: one second-order section ``[b0 b1 b2 1 a1 a2]`` with DF1 state."""

    def __init__(self, cutoff: float, q: float, fs: int | None = None) -> None:
        super().__init__()
        self.cutoff = cutoff
        self.q = q
        self.fs = fs
        self._sos: Tensor | None = None  # [1, 6] float64 on the host
        self._sos_device_cache: Tensor | None = None  # kept for API parity; the C ABI takes host coefficients
        self._state_x: Tensor | None = None  # [1, C, 2] float64
        self._state_y: Tensor | None = None

    @property
    def b(self) -> Tensor | None:
        return None if self._sos is None else self._sos[0, :3]

    @property
    def a(self) -> Tensor | None:
        if self._sos is None:
            return None
        return torch.tensor([1.0, float(self._sos[0, 4]), float(self._sos[0, 5])], dtype=torch.float64)

    def _set_coefficients(self, b0: float, b1: float, b2: float, a1: float, a2: float) -> None:
        self._sos = torch.tensor([[b0, b1, b2, 1.0, a1, a2]], dtype=torch.float64)
        self._sos_device_cache = None

    def _rbj(self) -> tuple[float, float, float]:
        """(cos w0, alpha, 1/(1+alpha)) of the RBJ cookbook for this cutoff / q / fs."""
        assert self.fs is not None
        _, cos_w0, alpha = self._compute_omega_alpha(self.cutoff, self.q, self.fs)
        return cos_w0, alpha, 1.0 / (1.0 + alpha)

    @staticmethod
    def _compute_omega_alpha(cutoff: float, q: float, fs: int) -> tuple[float, float, float]:
        w0 = 2.0 * math.pi * cutoff / fs
        sin_w0 = math.sin(w0)
        return sin_w0, math.cos(w0), sin_w0 / (2.0 * q)

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        if self.fs is None:
            raise ValueError("Sample rate (fs) must be set before filtering.")
        if self._sos is None:
            self.compute_coefficients()
        assert self._sos is not None
        y, self._state_x, self._state_y = run_sos_cascade(x, self._sos, self._state_x, self._state_y)
        return y

    def reset_state(self) -> None:
        """Forget the stream history (coefficients are kept, reference biquad.py:198-206)."""
        self._state_x = None
        self._state_y = None
        self._sos_device_cache = None

    def move_coeff(self, device) -> None:
        """No-op kept for callers of the pre-0.5 API (reference tests/test_cuda_kernels.py:43):
        coefficients travel as kernel parameters, there is nothing to move."""


class BiquadLPF(Biquad):
    def compute_coefficients(self) -> None:
        cos_w0, alpha, inv = self._rbj()
        mid = (1.0 - cos_w0) * inv
        self._set_coefficients(b0=mid / 2.0, b1=mid, b2=mid / 2.0, a1=-2.0 * cos_w0 * inv, a2=(1.0 - alpha) * inv)


class BiquadHPF(Biquad):
    def compute_coefficients(self) -> None:
        cos_w0, alpha, inv = self._rbj()
        mid = (1.0 + cos_w0) * inv
        self._set_coefficients(b0=mid / 2.0, b1=-mid, b2=mid / 2.0, a1=-2.0 * cos_w0 * inv, a2=(1.0 - alpha) * inv)


class BiquadNotch(Biquad):
    def compute_coefficients(self) -> None:
        cos_w0, alpha, inv = self._rbj()
        mid = -2.0 * cos_w0 * inv
        self._set_coefficients(b0=inv, b1=mid, b2=inv, a1=mid, a2=(1.0 - alpha) * inv)


class BiquadBPF(Biquad):
    """Constant 0 dB peak-gain band-pass."""

    def compute_coefficients(self) -> None:
        cos_w0, alpha, inv = self._rbj()
        self._set_coefficients(b0=alpha * inv, b1=0.0, b2=-alpha * inv, a1=-2.0 * cos_w0 * inv, a2=(1.0 - alpha) * inv)


class BiquadBPFPeak(Biquad):
    """Constant skirt-gain band-pass (peak gain = q)."""

    def compute_coefficients(self) -> None:
        cos_w0, alpha, inv = self._rbj()
        g = self.q * alpha * inv
        self._set_coefficients(b0=g, b1=0.0, b2=-g, a1=-2.0 * cos_w0 * inv, a2=(1.0 - alpha) * inv)


class BiquadAllPass(Biquad):
    def compute_coefficients(self) -> None:
        cos_w0, alpha, inv = self._rbj()
        lo = (1.0 - alpha) * inv
        mid = -2.0 * cos_w0 * inv
        self._set_coefficients(b0=lo, b1=mid, b2=1.0, a1=mid, a2=lo)
