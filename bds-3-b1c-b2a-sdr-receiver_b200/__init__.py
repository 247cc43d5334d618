"""B200-native acquisition / tracking correlator for the BDS-3 B1C / B2a SDR receiver.

Host-side mirror of the reference's MATLAB call surface (same function names,
argument meaning and struct layouts) on top of the C ABI of ``libbdsgpu.so``
(``include/bdsgpu.h``).  There is no CPU fallback: every compute call raises
``BdsError`` when the CUDA library or a B200 is missing.
"""
from ._lib import BdsError, lib, lib_path, device_ok, launch_count  # noqa: F401
from .settings import Settings  # noqa: F401
from . import b1c, b2a, codes, loopcoef, synth  # noqa: F401

__all__ = ["BdsError", "lib", "lib_path", "device_ok", "launch_count", "Settings", "b1c", "b2a", "codes",
           "loopcoef", "synth"]
