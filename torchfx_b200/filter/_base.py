"""Filter base classes: ``AbstractFilter`` and the parallel ``+`` combination.

Reference: src/torchfx/filter/__base.py -- ``AbstractFilter`` (:180-739:
``_has_computed_coeff`` :414-420, ``__add__`` / ``__radd__`` :530-739) and
``ParallelFilterCombination`` (:742-1026: children run on the same input, outputs summed,
:1019-1026).

B200 path: when every child is an SOS filter (``IIR`` / ``Biquad``) on a CUDA tensor the
whole combination is ONE launch of the band-per-lane filterbank kernel in SUM mode (x is
read once, y written once: 8 B/sample regardless of the number of children) instead of N
full passes plus N temporaries.
"""
from __future__ import annotations

import abc
from collections.abc import Sequence

import torch
from torch import Tensor

from ..effect import FX


class AbstractFilter(FX, abc.ABC):
    """Base of every filter; adds lazy coefficient design and the ``+`` operator."""

    @property
    def _has_computed_coeff(self) -> bool:
        if getattr(self, "_sos", None) is not None:
            return True
        if hasattr(self, "b") and hasattr(self, "a"):
            return self.b is not None and self.a is not None
        return False

    @abc.abstractmethod
    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)

    @abc.abstractmethod
    def compute_coefficients(self) -> None: ...

    def __add__(self, other: "AbstractFilter") -> "ParallelFilterCombination":
        assert isinstance(other, AbstractFilter), "Can only add AbstractFilter instances"
        return ParallelFilterCombination(self, other)

    def __radd__(self, other: "AbstractFilter") -> "ParallelFilterCombination":
        assert isinstance(other, AbstractFilter), "Can only add AbstractFilter instances"
        return ParallelFilterCombination(other, self)


class ParallelFilterCombination(AbstractFilter):
    """``f1 + f2 + ...``: every child filters the same input; the outputs are summed."""

    filters: Sequence[AbstractFilter]

    def __init__(self, *filters: AbstractFilter, fs: int | None = None) -> None:
        super().__init__()
        self.filters = filters
        self._bank = None  # lazily-built fused SOS bank (see _sosbank.py)
        # True: banks of 9..32 SOS children are added strictly in child order like the reference's loop
        # (:1019-1026) instead of per-warp partial sums (a rounding-level difference, the default is faster)
        self.strict_order = False
        self.fs = fs

    @property
    def _has_computed_coeff(self) -> bool:
        return all(f._has_computed_coeff for f in self.filters)

    @property
    def fs(self) -> int | None:
        return self._fs

    @fs.setter
    def fs(self, value: int | None) -> None:
        # children keep an fs they already have (reference __base.py:960-975,
        # tests/test_filter_base.py:88-130)
        self._fs = value
        if value is not None:
            for f in self.filters:
                if hasattr(f, "fs") and f.fs is None:
                    f.fs = value

    def compute_coefficients(self) -> None:
        for f in self.filters:
            f.compute_coefficients()

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        from ._sosbank import SosBank, bankable

        if x.is_cuda and bankable(self.filters):
            if self._bank is None:
                self._bank = SosBank(self.filters, mode="sum")
            from .. import _native as N

            self._bank.flags = (self._bank.flags & ~N.TFX_BANK_STRICT_ORDER) | (N.TFX_BANK_STRICT_ORDER if self.strict_order else 0)
            y = self._bank(x)
            if y is not None:
                return y
        total = torch.zeros_like(x)
        for f in self.filters:
            total += f.forward(x)
        return total
