"""FIR filters.

Reference: src/torchfx/filter/fir.py -- ``FIR`` (:284-579: taps stored FLIPPED as a
float32 ``[1, 1, K]`` buffer named ``kernel`` :516-518; causal, zero history, output
length T; ``conv_mode`` "fft" | "direct" | "auto" :510-514) and ``DesignableFIR``
(:582-1021: ``scipy.signal.firwin`` :1011-1018).  The reference evaluates the sum with
``torch.fft`` overlap-save (filter/_fftconv.py:107-141) or ``F.conv1d``; here both modes
call the library's own FIR kernels (``tfx_fir_f32``: shared-memory direct form for short
impulse responses, partitioned overlap-save block FFT for long ones; ``tfx_fir_f64``: direct
form in float64 for float64 signals) -- no torch.fft, no cuFFT, no conv1d.
"""
from __future__ import annotations

from collections.abc import Sequence

import torch
from numpy.typing import ArrayLike
from scipy.signal import firwin
from torch import Tensor

from .. import _native as N
from .. import _ops
from ..typing import WindowType
from ._base import AbstractFilter

_ALGO = {"fft": N.TFX_FIR_AUTO, "auto": N.TFX_FIR_AUTO, "direct": N.TFX_FIR_DIRECT}


def _auto_takes_overlap_save(K: int, samples: int) -> bool:
    """TFX_FIR_AUTO's rule (csrc/fir.cu pick_algo): direct form up to 56 taps, and up to 96 taps below 16 M samples."""
    return K > 96 or (K > 56 and samples >= 1 << 24)


def fir_plan(taps: Tensor) -> Tensor:
    """Overlap-save plan of a float32 impulse response on a CUDA device (``tfx_fir_plan_init``): the FFT twiddle tables
    and the spectra of the taps' partitions, computed once.  The reference re-runs ``rfft(kernel)`` inside every
    ``fft_conv1d`` call (filter/_fftconv.py:124)."""
    lib = N.load()
    h = taps.detach().reshape(-1).to(dtype=torch.float32).contiguous()
    if not h.is_cuda:
        raise ValueError("fir_plan needs the taps on a CUDA device")
    K = h.numel()
    nbytes = lib.tfx_fir_plan_bytes(K)
    plan = torch.empty(nbytes, dtype=torch.uint8, device=h.device)
    with torch.cuda.device(h.device):
        N.check(lib.tfx_fir_plan_init(h.data_ptr(), K, plan.data_ptr(), nbytes, torch.cuda.current_stream(h.device).cuda_stream))
    return plan


def fir_causal(x: Tensor, taps: Tensor, algo: int = N.TFX_FIR_AUTO, plan: Tensor | None = None) -> Tensor:
    """y[c, n] = sum_j taps[j] * x[c, n - j] over ``x`` ``[C, T]`` (zero history).  ``plan``: a ``fir_plan(taps)`` made
    on ``x``'s device; float32 CUDA signals then take the overlap-save kernel without transforming the taps again."""
    lib = N.load()
    if x.ndim != 2:
        raise ValueError(f"expected [C, T], got {tuple(x.shape)}")
    in_dtype = x.dtype
    cd = x.dtype if x.dtype == torch.float64 else torch.float32  # like the reference: the sum is evaluated in x.dtype
    xw = _ops._rows(x if x.dtype == cd else x.to(cd))
    C, T = xw.shape
    K = taps.numel()
    use_plan = plan is not None and xw.is_cuda and cd == torch.float32
    h = None if use_plan else taps.detach().reshape(-1).to(device=xw.device, dtype=cd).contiguous()
    y = torch.empty((C, T), dtype=cd, device=xw.device)
    ldx = xw.stride(0) if C > 1 else max(T, 1)
    if xw.is_cuda and cd == torch.float64:
        with torch.cuda.device(xw.device):
            N.check(lib.tfx_fir_f64(xw.data_ptr(), y.data_ptr(), C, T, ldx, max(T, 1), h.data_ptr(), K,
                                    torch.cuda.current_stream(xw.device).cuda_stream))
    elif use_plan:
        if plan.device != xw.device or plan.numel() < lib.tfx_fir_plan_bytes(K):
            raise ValueError("plan was made for another device or a longer impulse response")
        with torch.cuda.device(xw.device):
            nbytes = lib.tfx_fir_workspace_bytes(C, T, K, N.TFX_FIR_OLS)
            ws_ptr, ws_bytes = N.workspace(xw.device, nbytes)
            N.check(
                lib.tfx_fir_f32_planned(xw.data_ptr(), y.data_ptr(), C, T, ldx, max(T, 1), plan.data_ptr(), K, ws_ptr, ws_bytes,
                                        torch.cuda.current_stream(xw.device).cuda_stream)
            )
    elif xw.is_cuda:
        with torch.cuda.device(xw.device):
            nbytes = lib.tfx_fir_workspace_bytes(C, T, K, algo)
            ws_ptr, ws_bytes = N.workspace(xw.device, nbytes)
            N.check(
                lib.tfx_fir_f32(xw.data_ptr(), y.data_ptr(), C, T, ldx, max(T, 1), h.data_ptr(), K, algo, ws_ptr, ws_bytes,
                                torch.cuda.current_stream(xw.device).cuda_stream)
            )
    else:
        fn = lib.tfx_fir_cpu_f32 if cd == torch.float32 else lib.tfx_fir_cpu_f64
        N.check(fn(xw.data_ptr(), y.data_ptr(), C, T, ldx, max(T, 1), h.data_ptr(), K))
    return y if y.dtype == in_dtype else y.to(in_dtype)


class FIR(AbstractFilter):
    def __init__(self, b: ArrayLike, conv_mode: str = "fft") -> None:
        super().__init__()
        if conv_mode not in ("fft", "direct", "auto"):
            raise ValueError(f"conv_mode must be 'fft', 'direct', or 'auto', got {conv_mode!r}")
        self._conv_mode = conv_mode
        taps = torch.as_tensor(b, dtype=torch.float32).reshape(-1)
        self.a = [1.0]
        # same buffer name / layout as the reference so state_dicts stay interchangeable
        self.register_buffer("kernel", taps.flip(0)[None, None, :])
        self._plan: Tensor | None = None   # overlap-save plan of the current kernel on the device it was last used on
        self._plan_key: tuple | None = None

    def compute_coefficients(self) -> None:
        pass

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        shape = x.shape
        if x.ndim == 1:
            x2 = x.unsqueeze(0)
        elif x.ndim == 2:
            x2 = x
        elif x.ndim == 3:
            x2 = x.reshape(shape[0] * shape[1], shape[2])
        else:
            raise ValueError("Input must be of shape [T], [C, T], or [B, C, T]")
        taps = self.kernel[0, 0].flip(0)  # natural order b[0..K)
        algo = _ALGO[self._conv_mode]
        plan = None
        K = taps.numel()
        if x2.is_cuda and x2.dtype != torch.float64 and x2.numel() > 0 and (K > 1024 or (algo == N.TFX_FIR_AUTO and _auto_takes_overlap_save(K, x2.numel()))):
            # the overlap-save kernel will run: reuse the taps' spectra across calls (chunked callers); keyed by the
            # buffer's identity and version, so load_state_dict / in-place edits / .to() / a replaced buffer rebuild it
            # (the tensor OBJECT is kept, not its address: an address can be reused by a new buffer with version 0)
            src, ver, dev = self._plan_key if self._plan_key is not None else (None, -1, None)
            if src is not self.kernel or ver != self.kernel._version or dev != x2.device:
                self._plan = fir_plan(taps.to(x2.device))
                self._plan_key = (self.kernel, self.kernel._version, x2.device)
            plan = self._plan
        return fir_causal(x2, taps, algo, plan=plan).reshape(shape)


class DesignableFIR(FIR):
    def __init__(self, cutoff: float | Sequence[float], num_taps: int, fs: int | None = None, pass_zero: bool = True,
                 window: WindowType = "hamming", conv_mode: str = "fft") -> None:
        self.num_taps = num_taps
        self.cutoff = cutoff
        self.fs = fs
        self.pass_zero = pass_zero
        self.window = window
        self._conv_mode = conv_mode
        self.b: ArrayLike | None = None
        if fs is not None:
            self.compute_coefficients()

    def compute_coefficients(self) -> None:
        assert self.fs is not None, "Sampling frequency (fs) must be set."
        self.b = firwin(self.num_taps, self.cutoff, fs=self.fs, pass_zero=self.pass_zero, window=self.window, scale=True)
        super().__init__(self.b, conv_mode=self._conv_mode)
