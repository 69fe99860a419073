"""``Wave``: a ``[C, T]`` tensor plus its sample rate, with the deferred ``|`` pipeline.

Reference: src/torchfx/wave.py -- constructor/ys/to (:150-300), ``_materialize`` with the
IIR-run fuser (:207-239), ``_deferred`` (:241-257), ``__or__`` / fs propagation
(:578-703), ``merge`` / ``get_channel`` / ``duration`` (:705-900).  File I/O
(``from_file`` :406-470, ``save`` :472-576) reads and writes WAV through the package's own
RIFF codec (``_wavio``; soundfile is not installed here) and delegates other containers to
``soundfile`` when it is importable.
"""
from __future__ import annotations

import typing as tp
from collections.abc import Callable

import torch
from torch import Tensor, nn

from .effect import FX
from .filter._base import AbstractFilter
from .typing import Device


# Fold non-clamping Gain steps into the adjacent fused IIR run (see Wave._plan).
FOLD_GAIN = True


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
        (fresh instance => zero state per materialisation, reference wave.py:216-233).

        With ``FOLD_GAIN`` (module switch, default on) a non-clamping ``Gain`` between two
        such runs, or before / after one, does not cost a pass of its own (SURVEY.md 8f
        row 3): scaling is linear, so its factor is multiplied into the b-coefficients and
        the neighbouring fused runs become one cascade.  Folding never changes WHICH modules
        run statefully: a lone IIR between gains is run by the reference as the module
        itself, carrying its DF1 state from one materialisation to the next
        (wave.py:227-233), so it stays a step of its own here as well and gains that touch
        only such lone filters stay separate ``Gain`` steps."""
        from .effect import Gain
        from .filter.biquad import Biquad
        from .filter.fused import FusedSOSCascade
        from .filter.iir import IIR

        plan: list[nn.Module] = []
        part: list[nn.Module] = []  # gains and IIR runs of length >= 2 that may merge

        def close_part() -> None:
            filters = [m for m in part if isinstance(m, (IIR, Biquad))]
            if filters:
                g = 1.0
                for m in part:
                    if isinstance(m, Gain):
                        g *= m.linear_gain()
                plan.append(FusedSOSCascade(*filters, gain=g))
            else:
                plan.extend(part)
            part.clear()

        steps = self._pipeline
        i = 0
        while i < len(steps):
            step = steps[i]
            if isinstance(step, (IIR, Biquad)):
                j = i
                while j < len(steps) and isinstance(steps[j], (IIR, Biquad)):
                    j += 1
                if j - i >= 2:
                    part.extend(steps[i:j])
                else:  # a lone filter runs as itself (stateful) and separates what is around it
                    close_part()
                    plan.append(step)
                i = j
                continue
            if FOLD_GAIN and isinstance(step, Gain) and not step.clamp:
                part.append(step)
            else:
                close_part()
                plan.append(step)
            i += 1
        close_part()
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

    # ---- file I/O (reference wave.py:406-576) ------------------------------------------------
    @classmethod
    def from_file(cls, path, frame_offset: int = 0, num_frames: int = -1) -> "Wave":
        from . import _wavio

        stop = None if num_frames == -1 else frame_offset + num_frames
        if _wavio.is_wav_path(path):
            meta = _wavio.info(path)
            data_np, fs = _wavio.read(path, start=frame_offset, stop=stop, meta=meta)
            metadata = {"num_frames": meta.frames, "num_channels": meta.channels, "subtype": meta.subtype,
                        "format": meta.format}
        else:
            try:
                import soundfile as _sf
            except ImportError as e:  # pragma: no cover - depends on the image
                raise ImportError(f"reading {path} needs the 'soundfile' package (only WAV is built in)") from e
            data_np, fs = _sf.read(str(path), start=frame_offset, stop=stop, dtype="float32", always_2d=True)
            i = _sf.info(str(path))
            metadata = {"num_frames": i.frames, "num_channels": i.channels, "subtype": i.subtype, "format": i.format}
        return cls(torch.from_numpy(data_np.T.copy()), fs, metadata=metadata)

    def save(self, path, format: str | None = None, encoding: str | None = None,  # noqa: A002
             bits_per_sample: int | None = None) -> None:
        import os

        from . import _wavio

        parent = os.path.dirname(str(path))
        if parent:
            os.makedirs(parent, exist_ok=True)
        audio = self.ys.cpu()
        # encoding / bits_per_sample -> libsndfile subtype, as reference wave.py:548-565
        subtype: str | None = None
        if encoding is not None and bits_per_sample is not None:
            if encoding == "PCM_S":
                subtype = f"PCM_{bits_per_sample}"
            elif encoding == "PCM_U":
                subtype = "PCM_U8" if bits_per_sample == 8 else f"PCM_{bits_per_sample}"
            elif encoding == "PCM_F":
                subtype = "FLOAT" if bits_per_sample == 32 else "DOUBLE"
            else:
                subtype = f"{encoding}{bits_per_sample}"
        elif bits_per_sample is not None:
            subtype = f"PCM_{bits_per_sample}"
        elif encoding == "PCM_F":
            subtype = "FLOAT"
        frames = audio.numpy().T if audio.ndim == 2 else audio.numpy()
        if _wavio.is_wav_path(path, format):
            _wavio.write(path, frames, self.fs, subtype)
            return
        try:
            import soundfile as _sf
        except ImportError as e:  # pragma: no cover - depends on the image
            raise ImportError(f"writing {path} needs the 'soundfile' package (only WAV is built in)") from e
        ext = os.path.splitext(str(path))[1].lower()
        fmt = format or {".flac": "FLAC", ".ogg": "OGG"}.get(ext, "WAV")
        _sf.write(str(path), frames, self.fs, format=fmt, subtype=subtype)

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
