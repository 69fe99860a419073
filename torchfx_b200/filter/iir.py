"""IIR filter family: scipy-designed SOS cascades and cookbook biquads.

Reference: src/torchfx/filter/iir.py -- ``IIR`` base (:187-265), Butterworth / Chebyshev1 /
Chebyshev2 / Elliptic / LinkwitzRiley designed by ``scipy.signal.*(output="sos")``
(:380-385, :508-519, :639-650, :1950-1968, :2230-2242), convenience Hi*/Lo* classes
(defaults: ``LoButterworth`` / ``HiButterworth`` order 5, :868,918; ``order_scale="db"``
means ``order // 6``, :377), shelving / parametric / notch / all-pass cookbook sections
(:1098-1119, :1249-1270, :1441-1462, :1503-1512, :1635-1663, :1760-1790).

Coefficient design is host-side float64 and identical by construction (same scipy calls,
same closed forms); the arithmetic on the samples is the fused sm_100a cascade kernel.
"""
from __future__ import annotations

import abc
import math

import numpy as np
import torch
from scipy.signal import butter, cheby1, cheby2, ellip
from torch import Tensor

from ..typing import FilterOrderScale
from ._base import AbstractFilter
from ._cascade import run_sos_cascade
from .biquad import Biquad

NONE_FS_ERR = "Sample rate of the filter could not be None."


class IIR(AbstractFilter):
    """Base of the SOS-cascade IIR filters: lazy design, DF1 state carried across calls."""

    fs: int | None
    cutoff: float

    @abc.abstractmethod
    def __init__(self, fs: int | None = None) -> None:
        super().__init__()
        self.fs = fs
        self._sos: Tensor | None = None  # [K, 6] float64, host
        self._sos_device_cache: Tensor | None = None  # API parity only
        self._state_x: Tensor | None = None  # [K, C, 2] float64
        self._state_y: Tensor | None = None

    def _normalised_cutoff(self):
        assert self.fs is not None
        return np.asarray(self.cutoff, dtype=np.float64) / (0.5 * self.fs)

    def _store_sos(self, sos: np.ndarray) -> None:
        self._sos = torch.from_numpy(np.ascontiguousarray(sos, dtype=np.float64))

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        if self.fs is None:
            raise ValueError(NONE_FS_ERR)
        if self._sos is None:
            self.compute_coefficients()
        assert self._sos is not None
        y, self._state_x, self._state_y = run_sos_cascade(x, self._sos, self._state_x, self._state_y)
        return y

    def reset_state(self) -> None:
        """Forget the stream history AND the designed coefficients (reference iir.py:255-265)."""
        self._state_x = None
        self._state_y = None
        self._sos = None
        self._sos_device_cache = None

    def move_coeff(self, device) -> None:
        """No-op kept for callers of the pre-0.5 API (reference tests/test_cuda_kernels.py:43)."""


class Butterworth(IIR):
    def __init__(self, btype: str, cutoff: float, order: int = 4, order_scale: FilterOrderScale = "linear",
                 fs: int | None = None) -> None:
        super().__init__(fs)
        self.btype = btype
        self.cutoff = cutoff
        self.order = order if order_scale == "linear" else order // 6

    def compute_coefficients(self) -> None:
        self._store_sos(butter(self.order, self._normalised_cutoff(), btype=self.btype, output="sos"))


class Chebyshev1(IIR):
    def __init__(self, btype: str, cutoff: float, order: int = 4, ripple: float = 0.1, fs: int | None = None) -> None:
        super().__init__(fs)
        self.btype = btype
        self.cutoff = cutoff
        self.order = order
        self.ripple = ripple

    def compute_coefficients(self) -> None:
        self._store_sos(cheby1(self.order, self.ripple, self._normalised_cutoff(), btype=self.btype, output="sos"))


class Chebyshev2(IIR):
    def __init__(self, btype: str, cutoff: float, order: int = 4, ripple: float = 0.1, fs: int | None = None) -> None:
        super().__init__(fs)
        self.btype = btype
        self.cutoff = cutoff
        self.order = order
        self.ripple = ripple

    def compute_coefficients(self) -> None:
        self._store_sos(cheby2(self.order, self.ripple, self._normalised_cutoff(), btype=self.btype, output="sos"))


class Elliptic(IIR):
    def __init__(self, btype: str, cutoff: float, order: int = 4, passband_ripple: float = 0.1,
                 stopband_attenuation: float = 40, fs: int | None = None) -> None:
        super().__init__(fs)
        self.btype = btype
        self.cutoff = cutoff
        self.order = order
        self.passband_ripple = passband_ripple
        self.stopband_attenuation = stopband_attenuation

    def compute_coefficients(self) -> None:
        self._store_sos(
            ellip(self.order, self.passband_ripple, self.stopband_attenuation, self._normalised_cutoff(),
                  btype=self.btype, output="sos")
        )


class LinkwitzRiley(IIR):
    """Two identical Butterworth cascades of half the order (reference iir.py:2230-2242)."""

    def __init__(self, btype: str, cutoff: float, order: int = 4, order_scale: FilterOrderScale = "linear",
                 fs: int | None = None) -> None:
        super().__init__(fs)
        self.order = order if order_scale == "linear" else order // 6
        if order <= 0 or order % 2 != 0:
            raise ValueError("Linkwitz-Riley filter order must be a positive even integer.")
        self.btype = btype
        self.cutoff = cutoff

    def compute_coefficients(self) -> None:
        half = butter(self.order // 2, self._normalised_cutoff(), btype=self.btype, output="sos")
        self._store_sos(np.vstack([half, half]))


def _fixed_btype(base: type, btype: str, name: str, doc: str) -> type:
    """Hi*/Lo* convenience classes: ``base`` with ``btype`` bound as first argument."""

    def __init__(self, cutoff: float, *args, **kwargs) -> None:
        base.__init__(self, btype, cutoff, *args, **kwargs)

    return type(name, (base,), {"__init__": __init__, "__doc__": doc, "__module__": __name__})


class HiButterworth(Butterworth):
    def __init__(self, cutoff: float, order: int = 5, order_scale: FilterOrderScale = "linear", fs: int | None = None) -> None:
        super().__init__("highpass", cutoff, order, order_scale, fs)


class LoButterworth(Butterworth):
    def __init__(self, cutoff: float, order: int = 5, order_scale: FilterOrderScale = "linear", fs: int | None = None) -> None:
        super().__init__("lowpass", cutoff, order, order_scale, fs)


HiChebyshev1 = _fixed_btype(Chebyshev1, "highpass", "HiChebyshev1", "High-pass Chebyshev type I (cutoff, order=4, ripple=0.1, fs).")
LoChebyshev1 = _fixed_btype(Chebyshev1, "lowpass", "LoChebyshev1", "Low-pass Chebyshev type I (cutoff, order=4, ripple=0.1, fs).")
HiChebyshev2 = _fixed_btype(Chebyshev2, "highpass", "HiChebyshev2", "High-pass Chebyshev type II (cutoff, order=4, ripple=0.1, fs).")
LoChebyshev2 = _fixed_btype(Chebyshev2, "lowpass", "LoChebyshev2", "Low-pass Chebyshev type II (cutoff, order=4, ripple=0.1, fs).")
HiElliptic = _fixed_btype(Elliptic, "highpass", "HiElliptic", "High-pass elliptic (cutoff, order=4, passband_ripple=0.1, stopband_attenuation=40, fs).")
LoElliptic = _fixed_btype(Elliptic, "lowpass", "LoElliptic", "Low-pass elliptic (cutoff, order=4, passband_ripple=0.1, stopband_attenuation=40, fs).")
HiLinkwitzRiley = _fixed_btype(LinkwitzRiley, "highpass", "HiLinkwitzRiley", "High-pass Linkwitz-Riley (cutoff, order=4, order_scale, fs).")
LoLinkwitzRiley = _fixed_btype(LinkwitzRiley, "lowpass", "LoLinkwitzRiley", "Low-pass Linkwitz-Riley (cutoff, order=4, order_scale, fs).")


# ---- cookbook single-section filters ---------------------------------------------------------
class Shelving(Biquad):
    def __init__(self, cutoff: float, q: float, fs: int | None = None) -> None:
        super().__init__(cutoff=cutoff, q=q, fs=fs)

    @property
    def _omega(self) -> float:
        if self.fs is None:
            raise ValueError(NONE_FS_ERR)
        return 2.0 * math.pi * self.cutoff / self.fs

    @property
    def _alpha(self) -> float:
        return math.sin(self._omega) / (2.0 * self.q)

    def _shelf(self, sign: float) -> None:
        """RBJ shelf; ``sign`` = +1 high shelf, -1 low shelf (the two differ only in the sign
        of every cos term)."""
        A = self.gain
        c = sign * math.cos(self._omega)
        beta = 2.0 * math.sqrt(A) * self._alpha
        b0 = A * ((A + 1) + (A - 1) * c + beta)
        b1 = -2.0 * sign * A * ((A - 1) + (A + 1) * c)
        b2 = A * ((A + 1) + (A - 1) * c - beta)
        a0 = (A + 1) - (A - 1) * c + beta
        a1 = 2.0 * sign * ((A - 1) - (A + 1) * c)
        a2 = (A + 1) - (A - 1) * c - beta
        self._set_coefficients(b0=b0 / a0, b1=b1 / a0, b2=b2 / a0, a1=a1 / a0, a2=a2 / a0)


class HiShelving(Shelving):
    gain: float

    def __init__(self, cutoff: float, q: float, gain: float, gain_scale: FilterOrderScale = "linear",
                 fs: int | None = None) -> None:
        super().__init__(cutoff=cutoff, q=q, fs=fs)
        self.gain = gain if gain_scale == "linear" else 10 ** (gain / 20)

    def compute_coefficients(self) -> None:
        self._shelf(+1.0)


class LoShelving(Shelving):
    gain: float

    def __init__(self, cutoff: float, q: float, gain: float, gain_scale: FilterOrderScale = "linear",
                 fs: int | None = None) -> None:
        super().__init__(cutoff=cutoff, q=q, fs=fs)
        self.gain = gain if gain_scale == "linear" else 10 ** (gain / 20)

    def compute_coefficients(self) -> None:
        self._shelf(-1.0)


class ParametricEQ(Biquad):
    """Peaking EQ; ``gain`` in dB (reference iir.py:1438-1462)."""

    def __init__(self, frequency: float, q: float, gain: float, fs: int | None = None) -> None:
        super().__init__(cutoff=frequency, q=q, fs=fs)
        self.gain_db = gain
        self.gain = 10 ** (gain / 20)

    def compute_coefficients(self) -> None:
        cos_w0, alpha, _ = self._rbj()
        A = self.gain
        a0 = 1 + alpha / A
        self._set_coefficients(b0=(1 + alpha * A) / a0, b1=-2 * cos_w0 / a0, b2=(1 - alpha * A) / a0,
                               a1=-2 * cos_w0 / a0, a2=(1 - alpha / A) / a0)


class Peaking(ParametricEQ):
    def __init__(self, cutoff: float, q: float, gain: float, gain_scale: FilterOrderScale, fs: int | None = None) -> None:
        if gain_scale == "db":
            gain_db = gain
        else:
            gain_db = 20 * math.log10(gain) if gain > 0 else 0
        super().__init__(frequency=cutoff, q=q, gain=gain_db, fs=fs)


class Notch(Biquad):
    def __init__(self, cutoff: float, q: float, fs: int | None = None) -> None:
        super().__init__(cutoff=cutoff, q=q, fs=fs)

    def compute_coefficients(self) -> None:
        cos_w0, alpha, inv = self._rbj()
        mid = -2.0 * cos_w0 * inv
        self._set_coefficients(b0=inv, b1=mid, b2=inv, a1=mid, a2=(1.0 - alpha) * inv)


class AllPass(Biquad):
    def __init__(self, cutoff: float, q: float, fs: int | None = None) -> None:
        super().__init__(cutoff=cutoff, q=q, fs=fs)

    def compute_coefficients(self) -> None:
        cos_w0, alpha, inv = self._rbj()
        lo = (1.0 - alpha) * inv
        mid = -2.0 * cos_w0 * inv
        self._set_coefficients(b0=lo, b1=mid, b2=1.0, a1=mid, a2=lo)
