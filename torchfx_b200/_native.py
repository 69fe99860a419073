"""ctypes binding of ``libtorchfx_b200.so`` (C ABI: ``include/torchfx_b200.h``).

This module is the ONLY place that touches the shared library.  It fails loudly: if the
library is missing ``load()`` raises ``ImportError`` (there is no Python or torch
implementation to fall back to), and a device entry point called where no GPU is usable
raises ``RuntimeError`` carrying the library's message.

Reference counterpart: the pybind11 module ``torchfx.torchfx_ext``
(/root/reference/src/torchfx/_csrc/binding.cpp:83-96), hard-imported by
src/torchfx/_ops.py:25.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

import torch

LIB_NAME = "libtorchfx_b200.so"
# TFX_B200_LIB: developer override used by tools/ to A/B kernel-geometry variants.
LIB_PATH = os.environ.get("TFX_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", LIB_NAME)

# flags (mirror include/torchfx_b200.h)
TFX_OK = 0
TFX_EINVAL, TFX_ENODEVICE, TFX_ECUDA, TFX_EWORKSPACE, TFX_ENOMEM = -1, -2, -3, -4, -5
TFX_PREC_AUTO, TFX_PREC_F32, TFX_PREC_F64, TFX_NO_SPLIT, TFX_NO_TMA, TFX_FORCE_TMA, TFX_PACKED, TFX_NO_TILE = 0, 1, 2, 4, 8, 16, 32, 64
TFX_FORCE_TILE = 128
TFX_BANK_STRICT_ORDER = 256
TFX_BANK_STACK, TFX_BANK_SUM = 0, 1
TFX_FIR_AUTO, TFX_FIR_DIRECT, TFX_FIR_OLS = 0, 1, 2
TFX_SOS_MAX_K = 64
TFX_BANK_MAX_LANES = 64

_P = c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "tfx_version": (c_int, []),
    "tfx_last_error": (c_char_p, []),
    "tfx_device_count": (c_int, []),
    "tfx_kernel_launches": (c_uint64, []),
    "tfx_sos_cascade_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "tfx_sos_cascade_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int, _P, _P, c_uint32, _P, c_size_t, _P]),
    "tfx_sos_cascade_f64": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int, _P, _P, c_uint32, _P, c_size_t, _P]),
    "tfx_sos_auto_precision": (c_int, [_P, c_int, _P]),
    "tfx_sos_mixed_mask": (c_uint64, [_P, c_int, _P]),
    "tfx_sos_cascade_cpu_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int, _P, _P]),
    "tfx_sos_cascade_cpu_f64": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int, _P, _P]),
    "tfx_sos_cascade_host_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int, _P, _P, c_uint32, c_int64, c_int]),
    "tfx_filterbank_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "tfx_filterbank_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int, _P, _P, c_uint32, _P, c_size_t, _P]),
    "tfx_filterbank_f64": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int, _P, _P, c_uint32, _P, c_size_t, _P]),
    "tfx_plan_segmentation": (None, [c_int64, c_int64, c_int64, c_int64, c_int, _P, _P, _P]),
    "tfx_fir_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "tfx_fir_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int64, c_int, _P, c_size_t, _P]),
    "tfx_fir_f64": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int64, _P]),
    "tfx_fir_plan_bytes": (c_size_t, [c_int64]),
    "tfx_fir_plan_init": (c_int, [_P, c_int64, _P, c_size_t, _P]),
    "tfx_fir_f32_planned": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int64, _P, c_size_t, _P]),
    "tfx_fir_cpu_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int64]),
    "tfx_fir_cpu_f64": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int64]),
    "tfx_delay_line_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_double, c_double, _P]),
    "tfx_delay_line_f64": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_double, c_double, _P]),
    "tfx_delay_line_cpu_f32": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_double, c_double]),
    "tfx_delay_line_cpu_f64": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_double, c_double]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


class NativeError(RuntimeError):
    """A call into libtorchfx_b200.so failed (``code`` is the TFX_E* value)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"torchfx_b200 native error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C torchfx_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "torchfx_b200 has no pure-Python / torch fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load().tfx_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code: int) -> None:
    if code != TFX_OK:
        raise NativeError(code, last_error())


def kernel_launches() -> int:
    return int(load().tfx_kernel_launches())


def device_count() -> int:
    return int(load().tfx_device_count())


# ---- workspace cache (caller-owned device memory, as the C ABI requires) -----------------
_workspaces: dict[tuple[int, int], torch.Tensor] = {}


def workspace(device: torch.device, nbytes: int) -> tuple[int, int]:
    """Return (ptr, nbytes) of a cached uint8 device buffer of at least ``nbytes``.

    One buffer per (device, stream): launches on one stream are ordered, so reuse is safe.
    """
    if nbytes <= 0:
        return 0, 0
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf.data_ptr(), buf.numel()


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()
