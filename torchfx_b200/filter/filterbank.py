"""``LogFilterBank``: log-spaced constant-Q band-pass bank, output ``[n_bands, ...]``.

Reference: src/torchfx/filter/filterbank.py:47-185 -- ``n_bands`` ``BiquadBPF`` sections
at ``f_min * 2**(k * log2(f_max / f_min) / (n_bands - 1))``, q = 1.414 by default; the
reference loops over the bands in Python and ``torch.stack``s the results (:183-185).
Here a CUDA input takes ONE launch of the band-per-lane kernel: x is read once and every
band's output is written straight into its ``[n_bands, C, T]`` slot.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from ._base import AbstractFilter
from .biquad import BiquadBPF


class LogFilterBank(AbstractFilter):
    def __init__(self, n_bands: int, f_min: float = 20.0, f_max: float = 20000.0, q: float = 1.414,
                 fs: int | None = None) -> None:
        super().__init__()
        assert n_bands >= 2, "n_bands must be >= 2"
        assert 0 < f_min < f_max, "f_min must be positive and less than f_max"
        self.n_bands = n_bands
        self.f_min = f_min
        self.f_max = f_max
        self.q = q
        self._fs = fs
        span = math.log2(f_max / f_min)
        self._center_freqs = [f_min * 2.0 ** (k * span / (n_bands - 1)) for k in range(n_bands)]
        self.filters = [BiquadBPF(cutoff=fc, q=q, fs=fs) for fc in self._center_freqs]
        self._bank = None
        self.a: Tensor | None = None
        self.b: Tensor | None = None

    @property
    def fs(self) -> int | None:
        return self._fs

    @fs.setter
    def fs(self, value: int | None) -> None:
        self._fs = value
        if value is not None:
            for f in self.filters:
                f.fs = value

    @property
    def center_frequencies(self) -> list[float]:
        return list(self._center_freqs)

    def compute_coefficients(self) -> None:
        for f in self.filters:
            f.compute_coefficients()
        # sentinels so that _has_computed_coeff turns true (reference filterbank.py:166-167)
        self.a = torch.tensor([1.0])
        self.b = torch.tensor([1.0])

    def reset_state(self) -> None:
        for f in self.filters:
            f.reset_state()

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        if self._fs is None:
            raise ValueError("Sample rate (fs) must be set before filtering.")
        for f in self.filters:
            if f.fs is None:
                f.fs = self._fs
        if x.is_cuda:
            from ._sosbank import SosBank

            if self._bank is None:
                self._bank = SosBank(self.filters, mode="stack")
            y = self._bank(x)
            if y is not None:
                return y
        return torch.stack([f(x) for f in self.filters], dim=0)
