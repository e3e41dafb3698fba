"""pvd_b200 -- B200-native (sm_100a) implementation of PVD's volume-rendering hot path.

The native code lives in ``libpvd_b200.so`` (C ABI, see ``include/pvd_b200.h``); this package holds the loader and the
fused-path host logic.  The sibling packages ``raymarching``, ``gridencoder``, ``shencoder`` and ``tools`` mirror the
reference's operator API so its training scripts run unmodified with this directory on ``sys.path``.
"""
from . import _native  # noqa: F401

__all__ = ["_native"]
