"""Pyramid helpers: host-side mirror of ``src/zoom.py``.

The pyramid the reference's driver actually builds is ``skimage.transform.rescale``
(ica.py:333-336), available here as :func:`rescale` (CUDA, ``ica_rescale_host``).  The
reference's own ``zoom_out`` (zoom.py:29-60, the IPOL-style level) is dead code there -- it is never
called and raises on current scipy (scalar coordinates to ``map_coordinates``) -- so its mirror here
follows the algorithm it states and no reference run can pin it ("parity unpinned").
"""
from __future__ import annotations

import numpy as np

from . import _native
from . import constants as cts
from .transformation import _as_type


def zoom_size(nx, ny, factor):
    """``src/zoom.py:8-22``: ``int(np.round(n * factor))`` (round-half-to-even)."""
    return _native.zoom_size(int(nx), int(ny), float(factor))


def zoom_in_parameters(p, transformation_type, nx, ny, nxx, nyy):
    """``src/zoom.py:62-125``: up-scale the parameters from a (nx, ny) level to (nxx, nyy)."""
    try:
        t = _as_type(transformation_type)
    except ValueError:
        raise ValueError("Unsupported transformation type") from None
    return _native.zoom_in_parameters(np.asarray(p, dtype=np.float64), t.value, nx, ny, nxx, nyy)


def rescale(image, nu):
    """One pyramid level with the semantics of the reference's call
    ``rescale(I, nu, mode='constant', cval=0, order=3, anti_aliasing=True, channel_axis=2,
    preserve_range=True)`` (ica.py:333-336).  Returns float64."""
    return _native.rescale(image, float(nu)).astype(np.float64)


def zoom_out_operator(n, factor):
    """The 1-D linear operator of ``zoom_out`` along an axis of length ``n`` as a banded matrix:
    ``scipy.ndimage.gaussian_filter`` (sigma = ZOOM_SIGMA_ZERO * sqrt(1/factor^2 - 1), scipy's default ``reflect``
    extension, truncate 4) followed by cubic-spline ``map_coordinates(order=3, mode='nearest')`` at ``o / factor``.
    Built on the host by pushing the identity through scipy (the chain is linear); the GPU applies it.
    Returns ``(start [n_out] int32, weights [n_out, taps] float32)``."""
    from scipy import ndimage as ndi
    n_out = int(np.round(n * factor))
    sigma = cts.ZOOM_SIGMA_ZERO * np.sqrt(1.0 / (factor * factor) - 1.0)
    eye = np.eye(n)
    g = ndi.gaussian_filter1d(eye, sigma, axis=0)                    # column j = response to an impulse at j
    rows = np.arange(n_out, dtype=np.float64)[:, None] / factor + np.zeros((1, n))
    cols = np.zeros((n_out, 1)) + np.arange(n, dtype=np.float64)[None, :]
    a = ndi.map_coordinates(g, [rows, cols], order=3, mode="nearest")  # spline along the columns is the identity at nodes
    thr = 1e-9 * np.abs(a).max()
    nz = np.abs(a) > thr
    first = np.where(nz.any(1), nz.argmax(1), 0)
    last = np.where(nz.any(1), n - 1 - nz[:, ::-1].argmax(1), 0)
    taps = int(min(n, (last - first + 1).max()))
    start = np.clip(first - (taps - (last - first + 1)) // 2, 0, n - taps).astype(np.int32)
    weights = np.stack([a[o, start[o]:start[o] + taps] for o in range(n_out)]).astype(np.float32)
    return start, weights


def zoom_out(I, factor):
    """``src/zoom.py:29-60``: the IPOL-style pyramid level -- Gaussian pre-smoothing of every channel, then cubic-spline
    resampling at ``(i / factor, j / factor)``; output shape ``zoom_size``; no clipping.  The two 1-D operators are
    built on the host (:func:`zoom_out_operator`), the image is filtered on the GPU (``ica_apply_operators_host``).
    Returns float64."""
    img = np.asarray(I)
    if img.ndim != 3:
        raise ValueError("I must be (ny, nx, nz)")
    ny, nx, nz = img.shape
    ys, yw = zoom_out_operator(ny, float(factor))
    xs, xw = zoom_out_operator(nx, float(factor))
    if nz in (1, 3):
        return _native.apply_operators(img, ys, yw, xs, xw, clip=False).astype(np.float64)
    return np.concatenate([_native.apply_operators(img[:, :, c:c + 1], ys, yw, xs, xw, clip=False)
                           for c in range(nz)], axis=2).astype(np.float64)
