"""TEST INFRASTRUCTURE ONLY -- loads the unmodified reference for validation/baselines.

Two levels:

* ``load_ref_ext()``  -- the reference's compiled CPU extension
  (``oracle/_ref/torchfx_ext.so``, built by ``make -C oracle ref`` from the reference's
  own ``_csrc`` sources).  Travels to the GPU box; used by ``bench.py --impl reference``
  and the ``cpu_baseline`` leg.
* ``import_reference()`` -- the whole reference Python package, imported from
  ``/root/reference/src`` (exists only in the build container).  Used by
  ``oracle/make_golden.py`` and by tests that are skipped when the tree is absent.
  ``soundfile`` is stubbed: the reference imports it eagerly
  (src/torchfx/realtime/stream.py:28) but the filter path never touches it.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "torchfx_ext.so")
REF_SRC = os.environ.get("TORCHFX_REFERENCE_SRC", "/root/reference/src")


def have_ref_ext() -> bool:
    return os.path.exists(REF_SO)


def have_reference_tree() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "torchfx"))


def load_ref_ext():
    """Return the reference's pybind module (biquad_forward, sos_forward, delay_line_forward)."""
    if "torchfx_ext" in sys.modules:
        return sys.modules["torchfx_ext"]
    if not have_ref_ext():
        raise FileNotFoundError(f"{REF_SO} missing -- run `make -C oracle ref` in the build container")
    import torch  # noqa: F401  (libtorch must be loaded first)

    loader = importlib.machinery.ExtensionFileLoader("torchfx_ext", REF_SO)
    spec = importlib.util.spec_from_loader("torchfx_ext", loader, origin=REF_SO)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules["torchfx_ext"] = mod
    return mod


def import_reference():
    """Import the unmodified reference package ``torchfx`` (build container only)."""
    if "torchfx" in sys.modules and getattr(sys.modules["torchfx"], "__graft_ref__", False):
        return sys.modules["torchfx"]
    if not have_reference_tree():
        raise FileNotFoundError(f"reference tree not found at {REF_SRC}")
    ext = load_ref_ext()
    sys.modules["torchfx.torchfx_ext"] = ext
    if "soundfile" not in sys.modules:
        try:
            import soundfile  # noqa: F401
        except Exception:
            sys.modules["soundfile"] = types.ModuleType("soundfile")
    if "sounddevice" not in sys.modules:
        try:
            import sounddevice  # noqa: F401
        except Exception:
            pass
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import torchfx

    torchfx.__graft_ref__ = True
    return torchfx
