"""deeptreeattention_b200 -- B200-native (sm_100a) Hang2020 hot path for DeepTreeAttention.

``deeptreeattention_b200.Hang2020`` mirrors the reference's ``src.models.Hang2020`` module
namespace; ``_capi`` is the ctypes binding of the C-ABI library ``libdta_b200.so``.
"""
from . import _capi  # noqa: F401
from . import Hang2020  # noqa: F401

__all__ = ["Hang2020", "_capi"]
__version__ = "0.1.0"
