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
    Built in the native library in fp64 (``ica_zoom_out_operator``, csrc/ica_pyramid.cu: build_zoom_out_1d) -- no scipy
    in the product path; the GPU applies it.  Returns ``(start [n_out] int32, weights [n_out, taps] float32)``."""
    return _native.zoom_out_operator(int(n), float(factor))


def zoom_out(I, factor):
    """``src/zoom.py:29-60``: the IPOL-style pyramid level -- Gaussian pre-smoothing of every channel, then cubic-spline
    resampling at ``(i / factor, j / factor)``; output shape ``zoom_size``; no clipping.  Both 1-D operators are built
    in C++ and applied by the pyramid kernels (``ica_zoom_out_host``).  Returns float64."""
    img = np.asarray(I)
    if img.ndim != 3:
        raise ValueError("I must be (ny, nx, nz)")
    nz = img.shape[2]
    if nz in (1, 3):
        return _native.zoom_out(img, float(factor)).astype(np.float64)
    return np.concatenate([_native.zoom_out(img[:, :, c:c + 1], float(factor)) for c in range(nz)], axis=2).astype(np.float64)
