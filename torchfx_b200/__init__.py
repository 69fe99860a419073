"""torchfx_b200 -- B200-native filter engine behind the torchfx ``Wave | Filter`` surface.

Drop-in for the reference's filter path (matteospanio/torchfx: ``Wave`` / ``FX`` /
``torchfx.filter`` / ``torchfx._ops``); every ``forward`` dispatches through the C ABI of
``libtorchfx_b200.so`` (``include/torchfx_b200.h``) to hand-written sm_100a kernels.
See DESIGN.md for the path and its boundary, INTEGRATION.md for the reference-side stub.
"""
from . import _native, _ops, dist, effect, filter, realtime  # noqa: F401
from ._ops import get_default_precision, is_native_available, set_default_precision, torchfx_ext  # noqa: F401
from .chain import FilterChain  # noqa: F401
from .effect import FX, Gain, Reverb  # noqa: F401
from .realtime import StreamProcessor  # noqa: F401
from .wave import Wave  # noqa: F401

__version__ = "0.1.0"
