"""The drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
fails loudly without a GPU, and the product never touches oracle/.  CPU only."""
from __future__ import annotations

import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from torchfx_b200 import _native, _ops


def _header_symbols() -> set[str]:
    src = open(os.path.join(ROOT, "include", "torchfx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(tfx_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/torchfx_b200.h but not exported"
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)


def test_version_and_error_string():
    lib = _native.load()
    assert lib.tfx_version() == 100
    assert isinstance(_native.last_error(), str)
    assert _native.kernel_launches() >= 0


def test_reference_ops_surface():
    # reference tests/test_ops_dispatch.py:21-35
    assert _ops.PARALLEL_SCAN_THRESHOLD == 2048
    assert _ops.is_native_available() is True
    from torchfx_b200 import torchfx_ext

    for name in ("biquad_forward", "sos_forward", "delay_line_forward"):
        assert hasattr(torchfx_ext, name)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_device_entry_points_fail_loudly_without_gpu():
    lib = _native.load()
    sos = torch.tensor([[1.0, 0, 0, 1, 0, 0]], dtype=torch.float64)
    buf = torch.zeros(16, dtype=torch.float32)
    rc = lib.tfx_sos_cascade_f32(buf.data_ptr(), buf.data_ptr(), 1, 16, 16, 16, sos.data_ptr(), 1, None, None, 0, None, 0, None)
    assert rc == _native.TFX_ENODEVICE
    assert "no CPU fallback" in _native.last_error()
    with pytest.raises(_native.NativeError):
        _native.check(rc)
    out = torch.zeros(16, dtype=torch.float32)
    assert lib.tfx_delay_line_f32(buf.data_ptr(), out.data_ptr(), 1, 16, 16, 16, 4, 0.5, 0.5, None) == _native.TFX_ENODEVICE
    assert _native.kernel_launches() == 0


def test_bad_arguments_are_rejected():
    lib = _native.load()
    buf = torch.zeros(16, dtype=torch.float32)
    sos = torch.tensor([[1.0, 0, 0, 1, 0, float("nan")]], dtype=torch.float64)
    assert lib.tfx_sos_cascade_cpu_f32(buf.data_ptr(), buf.data_ptr(), 1, 16, 16, 16, sos.data_ptr(), 1, None, None) == _native.TFX_EINVAL
    assert "not finite" in _native.last_error()
    good = torch.tensor([[1.0, 0, 0, 1, 0, 0]], dtype=torch.float64)
    assert lib.tfx_sos_cascade_cpu_f32(buf.data_ptr(), buf.data_ptr(), 1, 16, 16, 16, good.data_ptr(), 0, None, None) == _native.TFX_EINVAL
    assert lib.tfx_sos_cascade_cpu_f32(buf.data_ptr(), buf.data_ptr(), 1, 16, 8, 16, good.data_ptr(), 1, None, None) == _native.TFX_EINVAL
    st = torch.zeros(2, dtype=torch.float64)
    assert lib.tfx_sos_cascade_cpu_f32(buf.data_ptr(), buf.data_ptr(), 1, 16, 16, 16, good.data_ptr(), 1, st.data_ptr(), None) == _native.TFX_EINVAL


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torchfx_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|torch\.fft\.\w+\(|^\s*import torchaudio|^\s*import triton|\blfilter\(", text, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_auto_precision_policy():
    import scipy.signal as sps

    lib = _native.load()
    err = ctypes.c_double()
    sos = torch.from_numpy(sps.butter(8, 5000 / 24000, output="sos")).contiguous()
    assert lib.tfx_sos_auto_precision(sos.data_ptr(), 4, ctypes.byref(err)) == _native.TFX_PREC_F32
    assert 0 < err.value < 2e-6
    hp = torch.from_numpy(sps.butter(2, 20 / 24000, btype="highpass", output="sos")).contiguous()
    assert lib.tfx_sos_auto_precision(hp.data_ptr(), 1, ctypes.byref(err)) == _native.TFX_PREC_F64
    assert err.value > 2e-6


def test_mixed_precision_mask_policy():
    """cfg4 chain: only the ParametricEQ section needs the float64 recurrence."""
    import torchfx_b200 as fx

    lib = _native.load()
    err = ctypes.c_double()
    chain = [fx.filter.LoButterworth(5000, order=4, fs=48000), fx.filter.ParametricEQ(1000, q=2.0, gain=3.0, fs=48000),
             fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=48000)]
    for f in chain:
        f.compute_coefficients()
    sos = torch.cat([f._sos for f in chain]).contiguous()
    assert lib.tfx_sos_auto_precision(sos.data_ptr(), 4, None) == _native.TFX_PREC_F64
    mask = lib.tfx_sos_mixed_mask(sos.data_ptr(), 4, ctypes.byref(err))
    assert mask == 0b0100 and 0 < err.value <= 2e-6
    import scipy.signal as sps

    lp = torch.from_numpy(sps.butter(8, 5000 / 24000, output="sos")).contiguous()
    assert lib.tfx_sos_mixed_mask(lp.data_ptr(), 4, None) == 0  # float32 everywhere
    hp = torch.from_numpy(sps.butter(4, 20 / 24000, btype="highpass", output="sos")).contiguous()
    assert lib.tfx_sos_mixed_mask(hp.data_ptr(), 2, None) == 0b11  # no proper subset suffices
