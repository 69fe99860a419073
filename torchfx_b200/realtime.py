"""``StreamProcessor``: chunked processing of signals that live on the HOST (files, host
tensors) with state carried from chunk to chunk.

Reference: src/torchfx/realtime/stream.py -- constructor and validation (:70-83), effect
normalisation (:108-121), fs configuration (:123-161), ``process_file`` (:163-255),
``process_chunks`` (:257-347), properties (:349-362).  The reference reads a chunk, moves
it to the device, calls every effect, moves it back and writes it -- serially.  Here the
same surface sits on the library's streaming driver (``tfx_sos_cascade_host_f32``):

* when the effect list is a pure IIR run (IIR / Biquad filters and non-clamping gains)
  and ``overlap == 0``, the whole run is ONE fused cascade and large blocks of the signal
  stream through the device with the copy-in, the kernel and the copy-out of consecutive
  time chunks overlapped; the DF1 state is carried between blocks, so the result equals
  the reference's chunk-by-chunk loop (same state contract, filter/iir.py:135-144);
* anything else runs the reference's generic per-chunk loop (effects are ordinary
  modules; stateful filters keep their own state between calls).

``process_tensor`` / ``iter_chunks`` are the same two paths for data already in host
memory (what ``bench.py`` times as ``e2e`` is this call's native core).
"""
from __future__ import annotations

import os
from collections.abc import Generator, Iterable, Sequence

import torch
from torch import Tensor, nn

from . import _ops, _wavio
from .effect import FX, Gain
from .filter._base import AbstractFilter
from .filter.biquad import Biquad
from .filter.fused import FusedSOSCascade
from .filter.iir import IIR

# Frames handed to the native streaming driver per call on the fused path (the driver cuts
# them further into ~256 MB device chunks).  Bounds host memory for file processing.
_BLOCK_SAMPLES = 1 << 26


class StreamProcessor:
    def __init__(self, effects: Sequence[FX] | nn.Sequential, chunk_size: int = 65536, overlap: int = 0,
                 device: str = "cpu") -> None:
        if not isinstance(chunk_size, int) or chunk_size <= 0:
            raise ValueError(f"chunk_size must be positive, got {chunk_size}")
        if overlap < 0:
            raise ValueError(f"Overlap must be non-negative, got {overlap}")
        if overlap >= chunk_size:
            raise ValueError(f"Overlap ({overlap}) must be less than chunk_size ({chunk_size})")
        self._effects: list[FX] = self._normalize_effects(effects)
        self._chunk_size = chunk_size
        self._overlap = overlap
        self._device = device
        self._fused: FusedSOSCascade | None = None
        self._fused_key: tuple | None = None
        self._fused_state: tuple[Tensor, Tensor] | None = None

    def __enter__(self) -> "StreamProcessor":
        return self

    def __exit__(self, *exc) -> None:
        pass

    @staticmethod
    def _normalize_effects(effects) -> list[FX]:
        out: list[FX] = []
        for e in effects:
            if not isinstance(e, FX):
                raise TypeError("All effects must inherit from FX when used in StreamProcessor")
            out.append(e)
        return out

    def _configure_effects(self, fs: int) -> None:
        """fs propagation, Nyquist validation and lazy coefficient design (reference
        stream.py:123-161)."""
        nyquist = fs / 2.0
        for e in self._effects:
            # validated before the design call so that the message names the filter (the reference
            # checks after it, stream.py:146-156, where scipy's own error has already fired)
            if isinstance(e, AbstractFilter) and hasattr(e, "cutoff"):
                cutoff = e.cutoff
                if isinstance(cutoff, (int, float)) and cutoff >= nyquist:
                    raise ValueError(
                        f"{type(e).__name__} cutoff ({cutoff} Hz) must be below the Nyquist frequency ({nyquist} Hz) "
                        f"for sample rate {fs} Hz. Reduce the cutoff or use a higher sample rate file."
                    )
            if hasattr(e, "fs") and e.fs != fs:
                e.fs = fs
                if isinstance(e, AbstractFilter):
                    e.compute_coefficients()
                    reset = getattr(e, "reset_state", None)
                    if callable(reset):
                        reset()
                    self.reset_state(effects=False)
            if isinstance(e, AbstractFilter) and not e._has_computed_coeff:
                e.compute_coefficients()

    # ---- the fused streaming path ---------------------------------------------------------
    def _fused_cascade(self) -> FusedSOSCascade | None:
        """The effect list as one cascade, or None when it is not a pure IIR run."""
        if self._overlap != 0 or not str(self._device).startswith("cuda"):
            return None
        filters = [e for e in self._effects if isinstance(e, (IIR, Biquad))]
        gains = [e for e in self._effects if isinstance(e, Gain) and not e.clamp]
        if not filters or len(filters) + len(gains) != len(self._effects):
            return None
        if any(getattr(f, "_sos", None) is None and getattr(f, "fs", None) is None for f in filters):
            return None
        g = 1.0
        for e in gains:
            g *= e.linear_gain()
        for f in filters:
            if f._sos is None:
                f.compute_coefficients()
        # content of every filter's coefficients is part of the key: a filter redesigned between calls
        # (same object, same fs) must not keep streaming through the old fused cascade
        key = (tuple(id(f) for f in filters), tuple(getattr(f, "fs", None) for f in filters), g,
               tuple(f._sos.detach().cpu().contiguous().numpy().tobytes() for f in filters))
        if self._fused is None or self._fused_key != key:
            self._fused = FusedSOSCascade(*filters, gain=g)
            self._fused_key = key
            self._fused_state = None
        return self._fused

    def reset_state(self, effects: bool = True) -> None:
        """Forget the carried state (a new signal starts)."""
        self._fused_state = None
        if effects:
            for e in self._effects:
                reset = getattr(e, "reset_state", None)
                if callable(reset):
                    reset()

    def _run_fused(self, fused: FusedSOSCascade, block: Tensor) -> Tensor:
        C = block.shape[0]
        K = fused._num_sections
        st = self._fused_state
        if st is None or st[0].shape[1] != C or st[0].shape[0] != K:
            st = (torch.zeros(K, C, 2, dtype=torch.float64), torch.zeros(K, C, 2, dtype=torch.float64))
            self._fused_state = st
        return _ops.sos_cascade_host_(block, fused._sos, st[0], st[1], device=self._device)

    def _run_generic(self, chunk: Tensor) -> Tensor:
        if self._device != "cpu":
            chunk = chunk.to(self._device)
        for e in self._effects:
            chunk = e(chunk)
        return chunk.cpu() if self._device != "cpu" else chunk

    # ---- host tensors ---------------------------------------------------------------------
    @torch.no_grad()
    def process_tensor(self, x: Tensor, fs: int | None = None) -> Tensor:
        """``x``: HOST ``[C, T]`` (or ``[T]``); returns the processed host tensor.  State is
        carried over to the next call (``reset_state()`` starts a new signal)."""
        if fs is not None:
            self._configure_effects(fs)
        squeeze = x.ndim == 1
        x2 = x.unsqueeze(0) if squeeze else x
        fused = self._fused_cascade()
        if fused is not None and x2.dtype == torch.float32 and not x2.is_cuda:
            y = self._run_fused(fused, x2)
        else:
            y = torch.cat(list(self.iter_chunks(self._split(x2))), dim=-1) if x2.shape[-1] else x2.clone()
        return y.squeeze(0) if squeeze else y

    def _split(self, x: Tensor) -> Generator[Tensor, None, None]:
        hop = self._chunk_size - self._overlap
        T = x.shape[-1]
        for off in range(0, T, hop):
            yield x[..., off: off + self._chunk_size]

    @torch.no_grad()
    def iter_chunks(self, chunks: Iterable[Tensor]) -> Generator[Tensor, None, None]:
        """Generic per-chunk loop over an iterable of ``[C, n]`` host tensors (reference
        stream.py:309-347 with the file reads factored out)."""
        first = True
        for chunk in chunks:
            y = self._run_generic(chunk)
            if self._overlap > 0 and not first:
                y = y[:, self._overlap:]
            first = False
            yield y

    # ---- files ----------------------------------------------------------------------------
    def _open(self, input_path):
        if not _wavio.is_wav_path(input_path):
            raise ValueError(f"{input_path}: only WAV input is built in (soundfile is not available in this image)")
        meta = _wavio.info(input_path)
        self._configure_effects(meta.samplerate)
        return meta

    def _file_blocks(self, input_path, meta, frames_per_read: int, hop: int) -> Generator[Tensor, None, None]:
        off = 0
        while off < meta.frames:
            data, _ = _wavio.read(input_path, off, off + min(frames_per_read, meta.frames - off), meta)
            yield torch.from_numpy(data.T.copy())
            off += hop

    @torch.no_grad()
    def process_chunks(self, input_path) -> Generator[Tensor, None, None]:
        """Yield processed ``[channels, n]`` host tensors (reference stream.py:257-347).  On the
        fused path the blocks are larger than ``chunk_size`` (the native driver does its own
        chunking); their concatenation is the same signal."""
        meta = self._open(input_path)
        fused = self._fused_cascade()
        if fused is not None:
            per = max(self._chunk_size, _BLOCK_SAMPLES // max(meta.channels, 1))
            for block in self._file_blocks(input_path, meta, per, per):
                yield self._run_fused(fused, block)
            return
        hop = self._chunk_size - self._overlap
        yield from self.iter_chunks(self._file_blocks(input_path, meta, self._chunk_size, hop))

    @torch.no_grad()
    def process_file(self, input_path, output_path, format: str | None = None,  # noqa: A002
                     subtype: str | None = None) -> None:
        """Chunked file -> file (reference stream.py:163-255; WAV output defaults to FLOAT)."""
        parent = os.path.dirname(str(output_path))
        if parent:
            os.makedirs(parent, exist_ok=True)
        if not _wavio.is_wav_path(output_path, format):
            raise ValueError(f"{output_path}: only WAV output is built in (soundfile is not available in this image)")
        meta = self._open(input_path)
        with _wavio.WavWriter(output_path, meta.samplerate, meta.channels, subtype or "FLOAT") as out:
            for y in self.process_chunks(input_path):
                out.write(y.numpy().T)

    # ---- properties (reference stream.py:349-362) ---------------------------------------------
    @property
    def chunk_size(self) -> int:
        return self._chunk_size

    @property
    def overlap(self) -> int:
        return self._overlap

    @property
    def effects(self) -> list[FX]:
        return list(self._effects)
