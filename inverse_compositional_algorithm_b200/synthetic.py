"""Synthetic image pairs with known ground-truth motion.

Host-side (numpy/scipy) generator shared by the tests and by ``bench.py``: a smooth random
texture, its centre crop as I2, and I1(x) = texture(x'(x; p_gt)) + noise, optionally with a
square occlusion.  This is input synthesis, not part of the registration path; the reference's
counterpart is ``transformation.transform_image`` (``src/transformation.py:266-318``), which
the notebooks use to fabricate test pairs.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

from .transformation import TransformType, project_points

SEED0 = 20240826


def random_motion(rng, transform_type: TransformType, height: int, width: int,
                  max_shift: float = 8.0, max_lin: float = 0.02) -> np.ndarray:
    """Ground-truth parameters in the reference's ordering for each model."""
    t = TransformType(transform_type)
    tx, ty = rng.uniform(-max_shift, max_shift, size=2)
    if t == TransformType.TRANSLATION:
        return np.array([tx, ty])
    if t == TransformType.EUCLIDEAN:
        return np.array([tx, ty, rng.uniform(-max_lin, max_lin)])
    if t == TransformType.SIMILARITY:
        a, b = rng.uniform(-max_lin, max_lin, size=2)
        return np.array([tx, ty, a, b])
    a00, a01, a10, a11 = rng.uniform(-max_lin, max_lin, size=4)
    if t == TransformType.AFFINITY:
        return np.array([tx, ty, a00, a01, a10, a11])
    h20, h21 = rng.uniform(-1.0, 1.0, size=2) * 1e-2 / max(height, width)
    return np.array([a00, a01, tx, a10, a11, ty, h20, h21])


def make_texture(rng, height: int, width: int, channels: int, margin: int,
                 sigma: float = 2.0) -> np.ndarray:
    tex = rng.standard_normal((height + 2 * margin, width + 2 * margin, channels))
    tex = ndi.gaussian_filter(tex, sigma=(sigma, sigma, 0.0), mode="wrap")
    lo, hi = tex.min(), tex.max()
    return ((tex - lo) * (255.0 / (hi - lo))).astype(np.float32)


def make_pair(seed: int, height: int, width: int, channels: int,
              transform_type: TransformType, *, max_shift: float = 8.0, max_lin: float = 0.02,
              noise_sigma: float = 1.0, occlusion: float = 0.0, margin: int = 64,
              p_gt=None):
    """Returns ``(I1, I2, p_gt)``: float32 ``(H, W, C)`` images and float64 parameters such that
    ``I1(x) ~= I2(x'(x; p_gt))``."""
    rng = np.random.default_rng(SEED0 + int(seed))
    tex = make_texture(rng, height, width, channels, margin)
    if p_gt is None:
        p_gt = random_motion(rng, transform_type, height, width, max_shift, max_lin)
    p_gt = np.asarray(p_gt, dtype=np.float64)
    i2 = np.ascontiguousarray(tex[margin:margin + height, margin:margin + width])
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    xp, yp = project_points(xx, yy, p_gt, transform_type)
    coords = np.stack([yp + margin, xp + margin])
    i1 = np.empty((height, width, channels), dtype=np.float64)
    for c in range(channels):
        i1[:, :, c] = ndi.map_coordinates(tex[:, :, c].astype(np.float64), coords, order=3,
                                          mode="nearest")
    if noise_sigma > 0:
        i1 += rng.normal(0.0, noise_sigma, size=i1.shape)
    if occlusion > 0:
        side = int(round(np.sqrt(occlusion * height * width)))
        side = min(side, height, width)
        y0 = int(rng.integers(0, height - side + 1))
        x0 = int(rng.integers(0, width - side + 1))
        i1[y0:y0 + side, x0:x0 + side] = rng.uniform(0.0, 255.0, size=(side, side, channels))
    np.clip(i1, 0.0, 255.0, out=i1)
    return i1.astype(np.float32), i2, p_gt
