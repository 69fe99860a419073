"""``SosBank``: N SOS filters applied to the same input in one launch (filterbank kernel).

Serves both parallel semantics of the reference: ``ParallelFilterCombination`` (outputs
summed, src/torchfx/filter/__base.py:1019-1026) and ``LogFilterBank`` (outputs stacked,
src/torchfx/filter/filterbank.py:183-185).  Children keep owning their DF1 state: after
every call each child's ``_state_x`` / ``_state_y`` is a view into the bank's
``[N, Kb, C, 2]`` tensors, so calling a child on its own afterwards continues its stream
exactly as in the reference.
"""
from __future__ import annotations

from collections.abc import Sequence

import torch
from torch import Tensor

from .. import _native as N
from .. import _ops

MAX_KB = 4  # sections per band the kernel is instantiated for


def bankable(filters: Sequence) -> bool:
    from .biquad import Biquad
    from .iir import IIR

    return len(filters) >= 1 and all(isinstance(f, (IIR, Biquad)) for f in filters)


class SosBank:
    def __init__(self, filters: Sequence, mode: str) -> None:
        assert mode in ("sum", "stack")
        self.filters = list(filters)
        self.mode = mode
        self._state_x: Tensor | None = None
        self._state_y: Tensor | None = None
        self._sos_key: tuple | None = None
        self._sos: Tensor | None = None
        self.flags = 0  # extra TFX_* kernel-selection flags (tests: TFX_NO_TILE, TFX_NO_SPLIT)

    def _gather_sos(self) -> Tensor | None:
        rows = []
        for f in self.filters:
            if f._sos is None:
                if f.fs is None:
                    raise ValueError("Sample rate (fs) must be set before filtering.")
                f.compute_coefficients()
            rows.append(f._sos)
        kb = max(r.shape[0] for r in rows)
        if kb > MAX_KB:
            return None
        # keyed on CONTENT: a redesign (new cutoff + compute_coefficients) may hand back a tensor at the same
        # address with the same version counter, and the rows are a few hundred bytes of host memory
        key = tuple(r.detach().cpu().contiguous().numpy().tobytes() for r in rows)
        if key != self._sos_key:
            ident = torch.tensor([1.0, 0.0, 0.0, 1.0, 0.0, 0.0], dtype=torch.float64)
            sos = ident.repeat(len(rows), kb, 1)
            for i, r in enumerate(rows):
                sos[i, : r.shape[0]] = r  # shorter cascades are padded with pass-through sections
            self._sos, self._sos_key = sos.contiguous(), key
        return self._sos

    def _gather_state(self, C: int, device: torch.device, kb: int) -> None:
        n = len(self.filters)
        fresh_x = torch.zeros(n, kb, C, 2, dtype=torch.float64, device=device)
        fresh_y = torch.zeros_like(fresh_x)
        for i, f in enumerate(self.filters):
            sx, sy = f._state_x, f._state_y
            if sx is not None and sy is not None and sx.shape[1] == C:
                k = sx.shape[0]
                fresh_x[i, :k] = sx.to(device)
                fresh_y[i, :k] = sy.to(device)
        self._state_x, self._state_y = fresh_x, fresh_y

    def __call__(self, x: Tensor) -> Tensor | None:
        """Returns None when this bank cannot be served by the fused kernel (caller loops)."""
        sos = self._gather_sos()
        if sos is None or x.dtype not in (torch.float32, torch.float64):
            return None
        shape = x.shape
        if x.ndim == 1:
            x2 = x.unsqueeze(0)
        elif x.ndim == 2:
            x2 = x
        elif x.ndim == 3:
            x2 = x.reshape(shape[0] * shape[1], shape[2])
        else:
            raise ValueError("Input must be of shape [T], [C, T], or [B, C, T]")
        x2 = _ops._rows(x2)
        C, T = x2.shape
        n, kb = sos.shape[0], sos.shape[1]
        # children may have been run (or reset) on their own since the last call
        views_ok = self._state_x is not None and self._state_x.shape[2] == C and self._state_x.device == x2.device and all(
            f._state_x is not None and f._state_x.data_ptr() == self._state_x[i].data_ptr()
            for i, f in enumerate(self.filters)
        )
        if not views_ok:
            self._gather_state(C, x2.device, kb)
        lib = N.load()
        if self.mode == "sum":
            y = torch.empty_like(x2)
            ldb = 0
        else:
            y = torch.empty((n, C, T), dtype=x2.dtype, device=x2.device)
            ldb = C * T
        suffix = "f32" if x2.dtype == torch.float32 else "f64"
        ldx = x2.stride(0) if C > 1 else max(T, 1)
        with torch.cuda.device(x2.device):
            nbytes = lib.tfx_filterbank_workspace_bytes(C, T, n, kb)
            ws_ptr, ws_bytes = N.workspace(x2.device, nbytes)
            fn = getattr(lib, f"tfx_filterbank_{suffix}")
            N.check(
                fn(x2.data_ptr(), y.data_ptr(), C, T, ldx, max(T, 1), ldb, sos.data_ptr(), n, kb,
                   N.TFX_BANK_SUM if self.mode == "sum" else N.TFX_BANK_STACK,
                   self._state_x.data_ptr(), self._state_y.data_ptr(),
                   _ops._PRECISIONS[_ops.get_default_precision()] | self.flags, ws_ptr, ws_bytes,
                   torch.cuda.current_stream(x2.device).cuda_stream)
            )
        for i, f in enumerate(self.filters):
            k = f._sos.shape[0]
            f._state_x = self._state_x[i, :k]
            f._state_y = self._state_y[i, :k]
        if self.mode == "sum":
            return y.reshape(shape)
        return y.reshape((n,) + tuple(shape))
