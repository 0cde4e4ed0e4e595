"""Jacobian / Hessian: host-side mirror of ``src/derivatives.py``.

On the registration path none of these arrays is ever materialised: the fused CUDA kernel
recomputes gradient, Jacobian and steepest-descent images per pixel and reduces straight to
H and b.  The functions here expose the same quantities for callers of the helper API.
"""
from __future__ import annotations

import numpy as np

from . import _native
from .transformation import TransformType, _as_type


def jacobian(transform_type, nx, ny):
    """``src/derivatives.py:7-70``: ``J[y, x, 0:n] = dx'/dp``, ``J[y, x, n:2n] = dy'/dp`` at p = 0,
    integer pixel coordinates (closed forms; the device code uses the same monomial table,
    ``csrc/ica_transform.cuh: jacobian_monomials``)."""
    t = _as_type(transform_type)
    n = t.nparams()
    J = np.zeros((ny, nx, 2 * n), dtype=np.float64)
    y, x = np.mgrid[0:ny, 0:nx]
    one = np.ones((ny, nx))
    rows = {
        TransformType.TRANSLATION: ([one, 0], [0, one]),
        TransformType.EUCLIDEAN: ([one, 0, -y], [0, one, x]),
        TransformType.SIMILARITY: ([one, 0, x, -y], [0, one, y, x]),
        TransformType.AFFINITY: ([one, 0, x, y, 0, 0], [0, one, 0, 0, x, y]),
        TransformType.HOMOGRAPHY: ([x, y, one, 0, 0, 0, -x * x, -x * y],
                                   [0, 0, 0, x, y, one, -x * y, -y * y]),
    }[t]
    for k in range(n):
        J[:, :, k] = rows[0][k]
        J[:, :, n + k] = rows[1][k]
    return J


def hessian_and_b(I1, I2, p, transform_type, robust_type=0, lambda_=0.0, nanifoutside=True,
                  delta=10):
    """One evaluation of the fused per-iteration kernel: the (rho'-weighted) Hessian
    (``src/derivatives.py:73-107``) and the vector b (``src/image_optimisation.py:82-143``) for the
    current parameters ``p``."""
    from .image_optimisation import _as_robust
    t = _as_type(transform_type)
    return _native.hessian_b(I1, I2, t.value, p, _as_robust(robust_type).value, lambda_, delta,
                             nanifoutside is True)


def hessian(DIJ):
    """``src/derivatives.py:73-88``: ``H = sum_{y,x,c} DIJ_c (x) DIJ_c`` on a materialised DIJ (non-finite entries
    zero-filled element-wise)."""
    return _native.dij_reduce(DIJ)


def hessian_robust(DIJ, rho, nparams):
    """``src/derivatives.py:91-107``: ``H = sum_{y,x} rho[y,x] sum_c DIJ_c (x) DIJ_c``."""
    return _native.dij_reduce(DIJ, rho=rho)


def inverse_hessian(H, nparams):
    """``src/derivatives.py:110-130``: LU inverse (partial pivoting); zero matrix when singular."""
    return _native.inverse_hessian(np.asarray(H, dtype=np.float64)[:nparams, :nparams])
