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
