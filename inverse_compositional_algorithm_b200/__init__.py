"""B200-native drop-in for the ``src/`` API of mfournigault/inverse_compositional_algorithm.

Python host code (this package) mirrors the reference's modules and function signatures;
all arithmetic on the registration path runs in hand-written sm_100a CUDA kernels behind a
C-ABI shared library (``include/ica_b200.h``), loaded with ctypes.  There is no CPU fallback:
importing ``_native`` fails loudly when the library is missing.
"""
from .transformation import TransformType  # noqa: F401
from .image_optimisation import RobustErrorFunctionType  # noqa: F401

__all__ = [
    "bicubic_interpolation",
    "configuration_handler",
    "constants",
    "derivatives",
    "image_optimisation",
    "inverse_compositional_algorithm",
    "transformation",
    "zoom",
]
