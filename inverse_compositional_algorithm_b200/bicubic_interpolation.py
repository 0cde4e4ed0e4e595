"""Bicubic warp: host-side mirror of ``src/bicubic_interpolation.py``.

``bicubic_interpolation_skimage`` is the warp the reference's drivers call (ica.py:111, 227); its
arithmetic is the stand-alone CUDA warp kernel (``ica_warp_host``), the same device code the
fused per-iteration kernel uses.
"""
from __future__ import annotations

import numpy as np

from . import _native
from .transformation import _as_type


def bicubic_interpolation_skimage(image, params, transformation_type, nanifoutside, delta):
    """``src/bicubic_interpolation.py:154-206``: order-3 (Catmull-Rom) warp of ``image`` by
    ``params2matrix(params)``; NaN where the 4x4 footprint leaves the image; result clipped to the
    image's [min, max]; identity when every ``|p_i| < 1e-10``.  ``nanifoutside`` and ``delta`` are
    accepted and ignored, exactly like the reference (its masking lines are commented out,
    bi.py:200-204).  Returns float64 like the reference."""
    t = _as_type(transformation_type)
    params = np.asarray(params, dtype=np.float64)
    if all(abs(v) < 1e-10 for v in params):
        m = np.eye(3)
    else:
        m = _native.params2matrix(params, t.value)
    img = np.asarray(image)
    out = _native.warp(img, m).astype(np.float64)
    return out[:, :, 0] if img.ndim == 2 else out


# ---- the IPOL-style warp and its scalar helpers (secondary API surface of the reference; not on the drivers' path)
def neumann_bc(x, nx):
    """``src/bicubic_interpolation.py:8-25``: clamp an index to ``[0, nx-1]``."""
    return 0 if x < 0 else (nx - 1 if x >= nx else x)


def cubic_interpolation(v, x):
    """``src/bicubic_interpolation.py:28-41``: Catmull-Rom through four samples at parameter ``x``."""
    return v[1] + 0.5 * x * (v[2] - v[0] + x * (2.0 * v[0] - 5.0 * v[1] + 4.0 * v[2] - v[3]
                                                + x * (3.0 * (v[1] - v[2]) + v[3] - v[0])))


def bicubic_interpolation_array(p, x, y):
    """``src/bicubic_interpolation.py:44-63``: 4x4 samples ``p[a][b]``, interpolated along ``b`` at ``y``, then at ``x``."""
    return cubic_interpolation([cubic_interpolation(p[a], y) for a in range(4)], x)


def bicubic_interpolation_image(input, params, nparams, nanifoutside, delta):
    """``src/bicubic_interpolation.py:121-152``: warp of the whole image, IPOL conventions -- the model is selected by
    ``nparams`` (2/3/4/6/8), NaN (``nanifoutside``) or 0 where the projected point is within ``delta`` of the border,
    Catmull-Rom with clamped neighbours, no clipping.  Computed on the GPU (``ica_warp_ipol_host``), float64."""
    if nparams not in (2, 3, 4, 6, 8):
        raise ValueError("Invalid transformation type")
    return _native.warp_ipol(np.asarray(input, dtype=np.float64), params, nparams, bool(nanifoutside), int(delta))


def bicubic_interpolation_point(input, uu, vv, nx, ny, nz, k):
    """``src/bicubic_interpolation.py:66-118``: the IPOL-style interpolation of channel ``k`` at ONE point (scalar host
    logic; whole images go through :func:`bicubic_interpolation_image` on the GPU)."""
    sx = -1 if uu < 0 else 1
    sy = -1 if vv < 0 else 1
    x, y = neumann_bc(int(uu), nx), neumann_bc(int(vv), ny)
    xs = [neumann_bc(int(uu) - sx, nx), x, neumann_bc(int(uu) + sx, nx), neumann_bc(int(uu) + 2 * sx, nx)]
    ys = [neumann_bc(int(vv) - sy, ny), y, neumann_bc(int(vv) + sy, ny), neumann_bc(int(vv) + 2 * sy, ny)]
    pol = [[float(input[ys[b], xs[a], k]) for b in range(4)] for a in range(4)]
    return bicubic_interpolation_array(pol, uu - x, vv - y)
