"""``Wave``: a ``[C, T]`` tensor plus its sample rate, with the deferred ``|`` pipeline.

Reference: src/torchfx/wave.py -- constructor/ys/to (:150-300), ``_materialize`` with the
IIR-run fuser (:207-239), ``_deferred`` (:241-257), ``__or__`` / fs propagation
(:578-703), ``merge`` / ``get_channel`` / ``duration`` (:705-900).  File I/O
(``from_file`` / ``save``, soundfile) is out of scope.
"""
from __future__ import annotations

import typing as tp
from collections.abc import Callable

import torch
from torch import Tensor, nn

from .effect import FX
from .filter._base import AbstractFilter
from .typing import Device


class Wave:
    fs: int

    def __init__(self, ys, fs: int, device: Device = "cpu", metadata: dict[str, tp.Any] | None = None) -> None:
        self.fs = fs
        self._pipeline: list[nn.Module] = []
        self._ys = ys if isinstance(ys, Tensor) else torch.as_tensor(ys, dtype=torch.float32)
        self.metadata = dict(metadata) if metadata else {}
        self._device: Device = device
        self.to(device)

    # ---- lazy data ------------------------------------------------------------------------
    @property
    def ys(self) -> Tensor:
        self._materialize()
        return self._ys

    @ys.setter
    def ys(self, value: Tensor) -> None:
        self._ys = value
        self._pipeline = []

    def _plan(self) -> list[nn.Module]:
        """Group consecutive IIR/Biquad steps; a run of >= 2 becomes ONE FusedSOSCascade
        (fresh instance => zero state per materialisation, reference wave.py:216-233)."""
        from .filter.biquad import Biquad
        from .filter.fused import FusedSOSCascade
        from .filter.iir import IIR

        plan: list[nn.Module] = []
        run: list[nn.Module] = []

        def close_run() -> None:
            if len(run) >= 2:
                plan.append(FusedSOSCascade(*run))
            else:
                plan.extend(run)
            run.clear()

        for step in self._pipeline:
            if isinstance(step, (IIR, Biquad)):
                run.append(step)
            else:
                close_run()
                plan.append(step)
        close_run()
        return plan

    def _materialize(self) -> None:
        if not self._pipeline:
            return
        data = self._ys
        for module in self._plan():
            data = module(data)
        self._ys = data
        self._pipeline = []

    @classmethod
    def _deferred(cls, ys: Tensor, fs: int, device: Device, metadata: dict[str, tp.Any], pipeline: list[nn.Module]) -> "Wave":
        w = object.__new__(cls)
        w._ys = ys
        w.fs = fs
        w._device = device
        w.metadata = metadata
        w._pipeline = pipeline
        return w

    # ---- device ---------------------------------------------------------------------------
    @property
    def device(self) -> Device:
        return self._device

    @device.setter
    def device(self, device: Device) -> None:
        self.to(device)

    def to(self, device: Device) -> "Wave":
        self._device = device
        self._materialize()
        self._ys = self._ys.to(device)
        return self

    def transform(self, func: Callable[..., Tensor], *args, **kwargs) -> "Wave":
        self._materialize()
        return Wave(func(self._ys, *args, **kwargs), self.fs)

    # ---- the pipe -------------------------------------------------------------------------
    def __or__(self, f: nn.Module) -> "Wave":
        if not isinstance(f, nn.Module):
            raise TypeError(f"Expected nn.Module, but got {type(f).__name__} instead.")
        for m in f.modules():
            if isinstance(m, FX):
                self._configure(m)
        steps = list(f.children()) if isinstance(f, nn.Sequential) else [f]
        return Wave._deferred(self._ys, self.fs, self._device, self.metadata, self._pipeline + steps)

    def _configure(self, f: FX) -> None:
        # reference wave.py:697-703: inherit the wave's fs, design coefficients once
        if hasattr(f, "fs") and f.fs is None:
            f.fs = self.fs
        if isinstance(f, AbstractFilter) and not f._has_computed_coeff:
            f.compute_coefficients()

    # ---- small accessors ------------------------------------------------------------------
    def __len__(self) -> int:
        return self.ys.shape[1]

    def channels(self) -> int:
        return self.ys.shape[0]

    def get_channel(self, index: int) -> "Wave":
        return Wave(self.ys[index], self.fs)

    def duration(self, unit: tp.Literal["sec", "ms"]) -> float:
        return len(self) / self.fs * (1000 if unit == "ms" else 1)

    @classmethod
    def merge(cls, waves: tp.Sequence["Wave"], split_channels: bool = False) -> "Wave":
        if not waves:
            raise ValueError("No waves to merge. Provide at least one wave.")
        fs = waves[0].fs
        for w in waves:
            if w.fs != fs:
                raise ValueError(
                    f"Sampling frequency mismatch: {w.fs} != {fs}. All waves must have the same sampling frequency."
                )
        if split_channels:
            return Wave(torch.cat([w.ys for w in waves], dim=0), fs)
        longest = max(len(w) for w in waves)
        first = waves[0].ys
        mix = torch.zeros((first.shape[0], longest), dtype=first.dtype, device=first.device)
        for w in waves:
            mix[:, : len(w)] += w.ys
        return Wave(mix, fs)
