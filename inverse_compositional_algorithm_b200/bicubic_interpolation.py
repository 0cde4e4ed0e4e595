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
