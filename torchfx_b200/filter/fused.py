"""``FusedSOSCascade``: a pipeline of IIR / Biquad filters collapsed into ONE kernel launch.

Reference: src/torchfx/filter/fused.py:19-132 (concatenate the SOS rows of N filters into
``[K_total, 6]``, own DF1 state, one native call; ``from_chain`` :87-107; ``move_coeff``
:109-111; ``reset_state`` :113-118).  On B200 the concatenated cascade runs in the fused
sm_100a kernel with all ``K_total`` sections' state in registers -- one 8 B/sample pass
however many filters were piped.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from ._cascade import run_sos_cascade
from .biquad import Biquad
from .iir import IIR


class FusedSOSCascade(nn.Module):
    def __init__(self, *filters: IIR | Biquad, gain: float = 1.0) -> None:
        """``gain`` (extension): a linear factor folded into the first
        section's b-coefficients, i.e. ``gain * H(z)`` at no extra pass over the signal."""
        super().__init__()
        if not filters:
            raise ValueError("FusedSOSCascade requires at least one IIR filter")
        rows: list[Tensor] = []
        fs: int | None = None
        for f in filters:
            if not hasattr(f, "_sos"):
                raise TypeError(f"Expected filter with SOS coefficients, got {type(f).__name__}")
            if f._sos is None:
                if f.fs is None:
                    raise ValueError(
                        f"Filter {type(f).__name__} has no sampling frequency set. Set fs before fusing."
                    )
                f.compute_coefficients()
            rows.append(f._sos)
            if f.fs is not None:
                if fs is None:
                    fs = f.fs
                elif f.fs != fs:
                    raise ValueError(f"Cannot fuse filters with different sample rates: {fs} vs {f.fs}")
        self._sos: Tensor = torch.cat(rows, dim=0).to(dtype=torch.float64, device="cpu")
        self.gain = float(gain)
        if self.gain != 1.0:
            self._sos = self._sos.clone()
            self._sos[0, :3] *= self.gain
        self._num_sections: int = self._sos.shape[0]
        self.fs: int | None = fs
        self._sos_device_cache: Tensor | None = None
        self._state_x: Tensor | None = None
        self._state_y: Tensor | None = None
        self._stateful: bool = False

    @classmethod
    def from_chain(cls, chain: nn.Sequential | nn.Module) -> "FusedSOSCascade":
        if isinstance(chain, nn.Sequential):
            members = [m for m in chain if isinstance(m, (IIR, Biquad))]
        elif isinstance(chain, (IIR, Biquad)):
            members = [chain]
        else:
            raise TypeError(f"Expected nn.Sequential or IIR/Biquad, got {type(chain).__name__}")
        if not members:
            raise ValueError("No IIR/Biquad filters found in chain to fuse")
        return cls(*members)

    def move_coeff(self, device) -> None:
        """Kept for API parity (reference fused.py:109-111).  Coefficients are kernel
        parameters taken from the host copy; nothing has to live on the device."""

    def reset_state(self) -> None:
        self._state_x = None
        self._state_y = None
        self._stateful = False
        self._sos_device_cache = None

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        y, self._state_x, self._state_y = run_sos_cascade(x, self._sos, self._state_x, self._state_y)
        self._stateful = True
        return y
