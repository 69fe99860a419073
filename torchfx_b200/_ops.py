"""Native dispatch layer -- drop-in for the reference's ``torchfx._ops``.

Reference: /root/reference/src/torchfx/_ops.py:34-191 (``PARALLEL_SCAN_THRESHOLD``,
``is_native_available``, ``biquad_forward``, ``parallel_iir_forward``,
``delay_line_forward``) and the pybind module behind it (``_csrc/binding.cpp:30-96``).

What differs from the reference, on purpose:

* CUDA tensors run hand-written sm_100a kernels through the C ABI of
  ``libtorchfx_b200.so`` (``include/torchfx_b200.h``) on torch's *current* stream (the
  reference launches on the legacy default stream, parallel_scan.cu:299-352).
* No float64 up-cast of the signal (reference ``_ops.py:142``): float32 in -> float32
  out, one 8 B/sample pass.  ``y`` is therefore returned in ``x.dtype`` rather than
  float64 (the reference's only caller casts back immediately, filter/iir.py:176).
* CPU tensors are served by the library's host twins (same device dispatch as
  binding.cpp:38,59,74).  A CUDA tensor is never computed on the CPU: if the kernels
  cannot run the call raises.
"""
from __future__ import annotations

import types
from contextlib import nullcontext

import torch
from torch import Tensor

from . import _native as N

# Public knob pinned by the reference's tests (tests/test_ops_dispatch.py:21-23).  Unused
# here: every length takes the same fused kernel.
PARALLEL_SCAN_THRESHOLD = 2048

_PRECISIONS = {"auto": N.TFX_PREC_AUTO, "f32": N.TFX_PREC_F32, "f64": N.TFX_PREC_F64}
_default_precision = "auto"


def set_default_precision(mode: str) -> None:
    """Recurrence arithmetic for float32 signals on CUDA: ``"auto"`` (per-filter policy,
    see ``tfx_sos_auto_precision``), ``"f32"`` or ``"f64"``."""
    global _default_precision
    if mode not in _PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {mode!r}")
    _default_precision = mode


def get_default_precision() -> str:
    return _default_precision


def is_native_available() -> bool:
    """True when ``libtorchfx_b200.so`` loads (reference ``_ops.py:37-54``)."""
    try:
        N.load()
    except (ImportError, OSError, AttributeError):
        return False
    return True


def _device_guard(t: Tensor):
    return torch.cuda.device(t.device) if t.is_cuda else nullcontext()


def _stream_ptr(t: Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _canon_sos(sos: Tensor) -> Tensor:
    """[K, 6] float64 contiguous on the HOST (the C ABI takes host coefficients)."""
    if sos.ndim != 2 or sos.shape[1] != 6:
        raise ValueError(f"sos must have shape [K, 6], got {tuple(sos.shape)}")
    s = sos.detach()
    if s.is_cuda:
        s = s.cpu()  # D2H sync; callers on the hot path pass the canonical CPU copy
    return s.to(dtype=torch.float64).contiguous()


def _rows(x: Tensor) -> Tensor:
    """Rows must be unit-stride; the row stride is passed to the library."""
    if x.stride(-1) != 1 and x.shape[-1] > 1:
        x = x.contiguous()
    if x.shape[0] > 1 and x.stride(0) < x.shape[1]:
        x = x.contiguous()
    return x


def _rows_aligned(x: Tensor, ld: int) -> bool:
    es = x.element_size()
    return x.data_ptr() % 16 == 0 and (ld * es) % 16 == 0


def _compute_dtype(x: Tensor) -> torch.dtype:
    return x.dtype if x.dtype in (torch.float32, torch.float64) else torch.float32


def sos_cascade_(
    x: Tensor,
    sos_cpu: Tensor,
    state_x: Tensor | None,
    state_y: Tensor | None,
    *,
    out: Tensor | None = None,
    precision: str | None = None,
    no_split: bool = False,
    no_tile: bool = False,
) -> Tensor:
    """Fused K-section cascade over ``x`` ``[C, T]``; updates ``state_x`` / ``state_y``
    (``[K, C, 2]`` float64 on ``x.device``) IN PLACE and returns ``y`` (``out`` if given;
    ``out`` may be ``x`` itself).
    """
    lib = N.load()
    if x.ndim != 2:
        raise ValueError(f"expected [C, T], got {tuple(x.shape)}")
    sos_cpu = _canon_sos(sos_cpu)
    K = sos_cpu.shape[0]
    in_dtype = x.dtype
    cd = _compute_dtype(x)
    xw = _rows(x if x.dtype == cd else x.to(cd))
    C, T = xw.shape
    if out is not None:
        if out.shape != xw.shape or out.dtype != cd or out.device != xw.device or (T > 1 and out.stride(-1) != 1):
            raise ValueError("out must match x in shape, dtype and device and have unit-stride rows")
        y = out
    else:
        y = torch.empty((C, T), dtype=cd, device=xw.device)
    for name, s in (("state_x", state_x), ("state_y", state_y)):
        if s is not None and (
            s.dtype != torch.float64 or s.device != xw.device or tuple(s.shape) != (K, C, 2) or not s.is_contiguous()
        ):
            raise ValueError(f"{name} must be a contiguous float64 [K={K}, C={C}, 2] tensor on {xw.device}")
    ldx = xw.stride(0) if C > 1 else max(T, 1)
    ldy = y.stride(0) if C > 1 else max(T, 1)
    suffix = "f32" if cd == torch.float32 else "f64"
    if xw.is_cuda and out is None and C > 1 and T >= 1024 and not _rows_aligned(xw, ldx):
        # Rows that do not start on 16-byte boundaries (odd T, sliced views) would take the
        # kernels' element-wise copy path (~3.5x slower).  Re-pitch once into a buffer whose row
        # stride is a multiple of 16 bytes (a plain strided copy: plumbing), filter in place there
        # and hand back the [C, T] view of it.
        vec = 16 // xw.element_size()
        ldp = (T + vec - 1) // vec * vec
        buf = torch.empty((C, ldp), dtype=cd, device=xw.device)
        y = buf[:, :T]
        y.copy_(xw)
        xw, ldx, ldy = y, ldp, ldp
    if xw.is_cuda:
        flags = (_PRECISIONS[precision or _default_precision] | (N.TFX_NO_SPLIT if no_split else 0)
                 | (N.TFX_NO_TILE if no_tile else 0))
        with _device_guard(xw):
            nbytes = lib.tfx_sos_cascade_workspace_bytes(C, T, K)
            ws_ptr, ws_bytes = N.workspace(xw.device, nbytes)
            fn = getattr(lib, f"tfx_sos_cascade_{suffix}")
            N.check(
                fn(xw.data_ptr(), y.data_ptr(), C, T, ldx, ldy, sos_cpu.data_ptr(), K, N.ptr(state_x), N.ptr(state_y),
                   flags, ws_ptr, ws_bytes, _stream_ptr(xw))
            )
    else:
        fn = getattr(lib, f"tfx_sos_cascade_cpu_{suffix}")
        N.check(fn(xw.data_ptr(), y.data_ptr(), C, T, ldx, ldy, sos_cpu.data_ptr(), K, N.ptr(state_x), N.ptr(state_y)))
    return y if y.dtype == in_dtype else y.to(in_dtype)


def sos_cascade_host_(
    x_host: Tensor,
    sos_cpu: Tensor,
    state_x: Tensor | None = None,
    state_y: Tensor | None = None,
    *,
    out: Tensor | None = None,
    device: int | str | torch.device = 0,
    chunk: int = 0,
    precision: str | None = None,
) -> Tensor:
    """HOST ``[C, T]`` float32 in, HOST float32 out, filtered on ``device`` by the library's
    streaming driver (``tfx_sos_cascade_host_f32``): time chunks of ``chunk`` samples (0 =
    library default) with the copy-in of chunk i+1, the kernel of chunk i and the copy-out
    of chunk i-1 overlapped, DF1 state carried on the device between chunks.  ``state_x`` /
    ``state_y`` are HOST ``[K, C, 2]`` float64 tensors updated in place (``None`` = zero
    state, discarded).  Pinned buffers (``x.pin_memory()``) run at full PCIe speed.

    Reference caller: the chunked file driver, src/torchfx/realtime/stream.py:160-347
    (per-chunk ``.to(device)`` / module call / ``.cpu()``, no overlap)."""
    lib = N.load()
    if x_host.ndim != 2:
        raise ValueError(f"expected [C, T], got {tuple(x_host.shape)}")
    if x_host.is_cuda or x_host.dtype != torch.float32:
        raise ValueError("sos_cascade_host_ takes a float32 HOST tensor (use sos_cascade_ for device tensors)")
    sos_cpu = _canon_sos(sos_cpu)
    K = sos_cpu.shape[0]
    xw = _rows(x_host)
    C, T = xw.shape
    if out is None:
        out = torch.empty((C, T), dtype=torch.float32, pin_memory=xw.is_pinned())
    elif out.shape != xw.shape or out.dtype != torch.float32 or out.is_cuda or (T > 1 and out.stride(-1) != 1):
        raise ValueError("out must be a float32 host tensor shaped like x with unit-stride rows")
    for name, s in (("state_x", state_x), ("state_y", state_y)):
        if s is not None and (s.dtype != torch.float64 or s.is_cuda or tuple(s.shape) != (K, C, 2) or not s.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous float64 host [K={K}, C={C}, 2] tensor")
    if (state_x is None) != (state_y is None):
        raise ValueError("state_x and state_y must both be given or both be None")
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    index = dev.index if dev.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    ldx = xw.stride(0) if C > 1 else max(T, 1)
    ldy = out.stride(0) if C > 1 else max(T, 1)
    N.check(lib.tfx_sos_cascade_host_f32(xw.data_ptr(), out.data_ptr(), C, T, ldx, ldy, sos_cpu.data_ptr(), K,
                                         N.ptr(state_x), N.ptr(state_y), _PRECISIONS[precision or _default_precision],
                                         int(chunk), index))
    return out


def _state_like(state: Tensor | None, shape: tuple[int, ...], device: torch.device) -> Tensor:
    """Fresh float64 state on ``device`` (zeros when ``None``), never aliasing the input:
    the reference's wrappers are functional (iir_cpu.cpp:72-73 clones)."""
    if state is None:
        return torch.zeros(shape, device=device, dtype=torch.float64)
    if tuple(state.shape) != shape:
        raise RuntimeError(f"state has shape {tuple(state.shape)}, expected {shape}")
    return state.detach().to(device=device, dtype=torch.float64, copy=True).contiguous()


def _as_2d(x: Tensor) -> tuple[Tensor, bool]:
    if x.ndim == 1:
        return x.unsqueeze(0), True
    if x.ndim != 2:
        raise RuntimeError(f"input must be [C, T] (or [T]), got {tuple(x.shape)}")
    return x, False


def biquad_forward(
    x: Tensor,
    b: Tensor,
    a: Tensor,
    state_x: Tensor | None,
    state_y: Tensor | None,
    *,
    a1_f64: float | None = None,
    a2_f64: float | None = None,
) -> tuple[Tensor, Tensor, Tensor]:
    """Single biquad section; returns ``(y, new_state_x, new_state_y)`` with states ``[C, 2]``.

    Same signature as the reference's ``_ops.biquad_forward`` (``_ops.py:57-116``).
    """
    x2, squeeze = _as_2d(x)
    C = x2.shape[0]
    bb = b.detach().to("cpu", torch.float64)
    if a1_f64 is None or a2_f64 is None:
        aa = a.detach().to("cpu", torch.float64)
        a1_f64, a2_f64 = float(aa[1]), float(aa[2])
    sos = torch.tensor([[float(bb[0]), float(bb[1]), float(bb[2]), 1.0, a1_f64, a2_f64]], dtype=torch.float64)
    sx = _state_like(state_x, (C, 2), x2.device).unsqueeze(0)
    sy = _state_like(state_y, (C, 2), x2.device).unsqueeze(0)
    y = sos_cascade_(x2, sos, sx, sy)
    return (y.squeeze(0) if squeeze else y), sx[0], sy[0]


def parallel_iir_forward(
    x: Tensor,
    sos: Tensor,
    state_x: Tensor | None,
    state_y: Tensor | None,
    *,
    sos_cpu: Tensor | None = None,
) -> tuple[Tensor, Tensor, Tensor]:
    """K-section SOS cascade; returns ``(y, new_state_x, new_state_y)`` with states ``[K, C, 2]``.

    Same signature as the reference's ``_ops.parallel_iir_forward`` (``_ops.py:119-176``).
    """
    x2, squeeze = _as_2d(x)
    C = x2.shape[0]
    coeffs = sos_cpu if sos_cpu is not None else sos
    K = coeffs.shape[0]
    sx = _state_like(state_x, (K, C, 2), x2.device)
    sy = _state_like(state_y, (K, C, 2), x2.device)
    y = sos_cascade_(x2, coeffs, sx, sy)
    return (y.squeeze(0) if squeeze else y), sx, sy


def delay_line_forward(x: Tensor, delay_samples: int, decay: float, mix: float) -> Tensor:
    """y[n] = x[n] + mix*decay*x[n-delay] (reference ``_ops.py:179-191``, delay_cpu.cpp:17-85)."""
    lib = N.load()
    if x.ndim not in (1, 2):
        raise RuntimeError(f"delay_line_forward expects [T] or [C, T], got {tuple(x.shape)}")
    if x.shape[-1] <= delay_samples:
        return x  # reference returns the input itself (delay_cpu.cpp:62-64)
    x2, squeeze = _as_2d(x)
    cd = _compute_dtype(x2)
    xw = _rows(x2 if x2.dtype == cd else x2.to(cd))
    C, T = xw.shape
    y = torch.empty((C, T), dtype=cd, device=xw.device)
    ldx = xw.stride(0) if C > 1 else T
    suffix = "f32" if cd == torch.float32 else "f64"
    if xw.is_cuda:
        with _device_guard(xw):
            fn = getattr(lib, f"tfx_delay_line_{suffix}")
            N.check(fn(xw.data_ptr(), y.data_ptr(), C, T, ldx, T, int(delay_samples), float(decay), float(mix), _stream_ptr(xw)))
    else:
        fn = getattr(lib, f"tfx_delay_line_cpu_{suffix}")
        N.check(fn(xw.data_ptr(), y.data_ptr(), C, T, ldx, T, int(delay_samples), float(decay), float(mix)))
    if y.dtype != x.dtype:
        y = y.to(x.dtype)
    return y.squeeze(0) if squeeze else y


def sos_forward(x: Tensor, sos: Tensor, sos_cpu: Tensor, state_x: Tensor, state_y: Tensor):
    """``torchfx_ext.sos_forward`` positional signature (binding.cpp:52-66)."""
    return parallel_iir_forward(x, sos, state_x, state_y, sos_cpu=sos_cpu)


def _ext_biquad_forward(x: Tensor, b: Tensor, a1: float, a2: float, state_x: Tensor, state_y: Tensor):
    """``torchfx_ext.biquad_forward`` positional signature (binding.cpp:30-50)."""
    return biquad_forward(x, b, torch.tensor([1.0, a1, a2], dtype=torch.float64), state_x, state_y, a1_f64=a1, a2_f64=a2)


# Object importable as ``torchfx_ext`` with the three attributes the reference's tests pin
# (tests/test_ops_dispatch.py:29-35).
torchfx_ext = types.SimpleNamespace(
    biquad_forward=_ext_biquad_forward,
    sos_forward=sos_forward,
    delay_line_forward=delay_line_forward,
)
