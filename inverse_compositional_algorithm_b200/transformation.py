"""Transform algebra: host-side mirror of the reference's ``src/transformation.py``.

Same names, argument meaning and error behaviour as the reference.  The arithmetic that the
registration loop needs on the device (matrix from parameters, p <- p o dp^-1) lives in
``csrc/ica_transform.cuh`` and is compiled for both host and device; the functions here that
are on the hot path call that single implementation through the C-ABI, so Python, the solve
epilogue kernel and the parity tests all run the same code.
"""
from __future__ import annotations

from enum import Enum

import numpy as np


class TransformType(Enum):
    """``src/transformation.py:8-13`` (values cross the C-ABI unchanged)."""
    TRANSLATION = 1
    EUCLIDEAN = 2
    SIMILARITY = 3
    AFFINITY = 4
    HOMOGRAPHY = 5

    def nparams(self) -> int:
        """``src/transformation.py:15-32``"""
        return _NPARAMS[self]


_NPARAMS = {
    TransformType.TRANSLATION: 2,
    TransformType.EUCLIDEAN: 3,
    TransformType.SIMILARITY: 4,
    TransformType.AFFINITY: 6,
    TransformType.HOMOGRAPHY: 8,
}


def _as_type(transform_type) -> TransformType:
    if isinstance(transform_type, TransformType):
        return transform_type
    try:
        return TransformType(int(transform_type))
    except (ValueError, TypeError):
        raise ValueError("Unknown transform type") from None


def params2matrix(p, transform_type) -> np.ndarray:
    """``src/transformation.py:188-236``: 3x3 matrix of x'(x; p) (native, fp64)."""
    from . import _native
    t = _as_type(transform_type)
    return _native.params2matrix(np.asarray(p, dtype=np.float64), t.value)


def matrix2params(matrix, transform_type):
    """``src/transformation.py:238-263`` (returns a list, like the reference)."""
    t = _as_type(transform_type)
    m = np.asarray(matrix, dtype=np.float64)
    if t == TransformType.TRANSLATION:
        return [m[0, 2], m[1, 2]]
    if t == TransformType.EUCLIDEAN:
        return [m[0, 2], m[1, 2], np.arctan2(m[1, 0], m[0, 0])]
    if t == TransformType.SIMILARITY:
        return [m[0, 2], m[1, 2], m[0, 0] - 1, m[1, 0]]
    if t == TransformType.AFFINITY:
        return [m[0, 2], m[1, 2], m[0, 0] - 1, m[0, 1], m[1, 0], m[1, 1] - 1]
    return [m[0, 0] - 1, m[0, 1], m[0, 2], m[1, 0], m[1, 1] - 1, m[1, 2], m[2, 0], m[2, 1]]


def transform_image(image, transformation_type, gt):
    """``src/transformation.py:266-318``: the notebooks' test-data generator -- the image warped by the INVERSE of the
    model matrix with skimage's defaults (bilinear, zero outside, clipped to the input range), on the GPU.
    Like the reference, TRANSLATION uses the translation only and EUCLIDEAN is built from ``rotation=-gt[2]``
    (``transformation.py:309``); all-zero parameters mean the identity."""
    from . import _native
    t = _as_type(transformation_type)
    gt = [float(v) for v in gt]
    if all(abs(v) < 1e-10 for v in gt):
        m = np.eye(3)
    elif t == TransformType.EUCLIDEAN:
        m = params2matrix([gt[0], gt[1], -gt[2]], t)
    elif t == TransformType.TRANSLATION:
        m = params2matrix(gt[:2], t)
    else:
        m = params2matrix(gt, t)
    return _native.transform_image(image, np.linalg.inv(np.asarray(m, dtype=np.float64)))


def project_points(x, y, p, transform_type):
    """x'(x; p) for arrays of points (numpy; used for EPE and data synthesis)."""
    t = _as_type(transform_type)
    p = np.asarray(p, dtype=np.float64)
    if t == TransformType.TRANSLATION:
        return x + p[0], y + p[1]
    if t == TransformType.EUCLIDEAN:
        c, s = np.cos(p[2]), np.sin(p[2])
        return c * x - s * y + p[0], s * x + c * y + p[1]
    if t == TransformType.SIMILARITY:
        return (1 + p[2]) * x - p[3] * y + p[0], p[3] * x + (1 + p[2]) * y + p[1]
    if t == TransformType.AFFINITY:
        return (1 + p[2]) * x + p[3] * y + p[0], p[4] * x + (1 + p[5]) * y + p[1]
    d = p[6] * x + p[7] * y + 1
    return ((1 + p[0]) * x + p[1] * y + p[2]) / d, (p[3] * x + (1 + p[4]) * y + p[5]) / d


def project(x, y, p, nparams):
    """``src/transformation.py:144-186``: one point; the model is selected by the NUMBER of
    parameters (2/3/4/6/8), exactly like the reference."""
    by_n = {2: TransformType.TRANSLATION, 3: TransformType.EUCLIDEAN,
            4: TransformType.SIMILARITY, 6: TransformType.AFFINITY, 8: TransformType.HOMOGRAPHY}
    if nparams not in by_n:
        raise ValueError("Invalid transformation type")
    return project_points(x, y, p, by_n[nparams])


def update_transform(p, dp, transform_type):
    """``src/transformation.py:36-141``: p <- params(M(p) M(dp)^-1), IN PLACE on ``p`` when it
    is a float64 array (the reference mutates its argument), value also returned.  Runs the
    library's host/device-shared closed forms (reference formulas term by term)."""
    from . import _native
    t = _as_type(transform_type)
    out = _native.update_transform(np.asarray(p, dtype=np.float64),
                                   np.asarray(dp, dtype=np.float64), t.value)
    if isinstance(p, np.ndarray) and p.dtype == np.float64:
        p[...] = out
        return p
    return out


def end_point_error(pa, pb, transform_type, nx, ny):
    """Mean and max over the image domain of ||x'(x;pa) - x'(x;pb)||."""
    y, x = np.mgrid[0:ny, 0:nx].astype(np.float64)
    xa, ya = project_points(x, y, pa, transform_type)
    xb, yb = project_points(x, y, pb, transform_type)
    d = np.hypot(xa - xb, ya - yb)
    return float(d.mean()), float(d.max())
