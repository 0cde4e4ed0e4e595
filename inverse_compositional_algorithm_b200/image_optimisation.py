"""Robust error functions: host-side mirror of ``src/image_optimisation.py``.

The per-pixel work of this module (rho', the steepest-descent.residual vector b) is fused into
the per-iteration CUDA kernel (``csrc/ica_iterate.cu``); here live the Enum and thin helpers.
"""
from __future__ import annotations

from enum import Enum

import numpy as np


class RobustErrorFunctionType(Enum):
    """``src/image_optimisation.py:10-15`` (values cross the C-ABI unchanged; GERMAN_MCCLURE
    keeps the reference's spelling)."""
    QUADRATIC = 0
    TRUNCATED_QUADRATIC = 1
    GERMAN_MCCLURE = 2
    LORENTZIAN = 3
    CHARBONNIER = 4


def _as_robust(robust_type) -> RobustErrorFunctionType:
    if isinstance(robust_type, RobustErrorFunctionType):
        return robust_type
    try:
        return RobustErrorFunctionType(int(robust_type))
    except (ValueError, TypeError):
        raise ValueError("Unknown type for robust error function") from None


def parametric_solve(H_1, b, nparams):
    """``src/image_optimisation.py:146-155``: dp = H^-1 b, error = ||dp||_2 (n <= 8, host)."""
    dp = np.asarray(H_1, dtype=np.float64) @ np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.sum(dp ** 2))), dp


# ---- helpers on materialised arrays (the drivers never build them; kept for callers of the reference's helper API)
def rhop(t2, lambda_, type_):
    """``src/image_optimisation.py:17-53``: derivative of the robust error function, element-wise on the GPU
    (TRUNCATED_QUADRATIC element-wise; the reference's array branch raises)."""
    from . import _native
    return _native.rhop(t2, lambda_, _as_robust(type_).value)


def robust_error_function(DI, lambda_, type_):
    """``src/image_optimisation.py:56-79``: rho'(sum_c DI_c^2) per pixel, non-finite DI zero-filled."""
    from . import _native
    return _native.robust_error(DI, lambda_, _as_robust(type_).value)


def steepest_descent_images(Ix, Iy, J, nparams):
    """``src/image_optimisation.py:158-194``: ``DIJ[y,x,c,k] = Ix[y,x,c] J[y,x,k] + Iy[y,x,c] J[y,x,k+n]``."""
    from . import _native
    if np.shape(Ix) != np.shape(Iy):
        raise ValueError("Ix and Iy must have the same dimensions")
    return _native.steepest_descent(Ix, Iy, J, nparams)


def independent_vector(DIJ, DI, nparams):
    """``src/image_optimisation.py:82-110``: ``b = sum DIJ^T DI`` (non-finite factors zero-filled)."""
    from . import _native
    return _native.dij_reduce(DIJ, DI=DI)


def independent_vector_robust(DIJ, DI, rho, nparams):
    """``src/image_optimisation.py:113-143``: ``b = sum rho DIJ^T DI``."""
    from . import _native
    return _native.dij_reduce(DIJ, DI=DI, rho=rho)
