"""chomp_b200 -- B200-native halo-model / Limber / Hankel hot path of CHOMP.

Drop-in modules with the reference's names (``cosmology``, ``mass_function``,
``hod``, ``halo``, ``kernel``, ``correlation``, ``defaults``) plus the batched
interface (``engine.Engine``, ``engine.Survey``, ``design``).  All numerics run
in hand-written sm_100a CUDA kernels behind the C ABI of include/chomp_b200.h;
there is no CPU fallback.
"""
from . import defaults  # noqa: F401
from ._lib import ChompError, EXPORTED_SYMBOLS, LIB_PATH  # noqa: F401

__all__ = ["defaults", "ChompError", "EXPORTED_SYMBOLS", "LIB_PATH"]
