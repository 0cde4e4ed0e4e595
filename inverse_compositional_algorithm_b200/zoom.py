"""Pyramid helpers: host-side mirror of ``src/zoom.py``.

The pyramid the reference's driver actually builds is ``skimage.transform.rescale``
(ica.py:333-336), available here as :func:`rescale` (CUDA, ``ica_rescale_host``).  The
reference's own ``zoom_out`` (zoom.py:29-60) is dead code that raises on current scipy and is
not provided.
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
