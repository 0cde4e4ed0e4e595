"""CPU restatement of the two scikit-image calls on the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``inverse_compositional_algorithm_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may use it.

The reference (``/root/reference/src``) delegates its image warp and its
pyramid to scikit-image 0.24.0 (``requirements.txt:4``), which is NOT installed
in this image and cannot be installed (no network).  This module restates the
published behaviour of

* ``skimage.transform.warp(image, tform, order, mode="constant", cval, clip=True,
  preserve_range=True)`` as called at ``src/bicubic_interpolation.py:199``
  (order 3, cval NaN) and ``src/transformation.py:316`` (order 1, cval 0), and
* ``skimage.transform.rescale(image, s, mode='constant', cval=0, order=3,
  anti_aliasing=True, channel_axis=2, preserve_range=True)`` as called at
  ``src/inverse_compositional_algorithm.py:333-336``

on top of numpy and scipy.ndimage (scipy IS installed; skimage itself calls
``scipy.ndimage.gaussian_filter`` and ``scipy.ndimage.zoom`` for ``rescale``).
Semantics are those written up in ``SURVEY.md`` Appendix A.  The restatement is
pinned by the reference's own stored notebook outputs: running the unmodified
reference sources on top of it reproduces the 16-digit per-iteration
trajectories in ``test/inverse_compositional_algorithm_robust.ipynb`` and
``test/inverse_compositional_algorithm.ipynb`` (see ``tests/test_oracle_kat.py``
and ``tests/golden/notebook_trajectories.json``).
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi


# --------------------------------------------------------------------------- clip
def clip_like_skimage(input_image, output_image, mode, cval, clip=True):
    """skimage ``_clip_warp_output``: clamp to the input's [min, max] (NaN-aware),
    keeping ``cval`` pixels when cval lies outside that range but was produced."""
    if not clip:
        return output_image
    lo = np.min(input_image)
    if np.isnan(lo):
        fmin, fmax = np.nanmin, np.nanmax
        lo = fmin(input_image)
    else:
        fmin, fmax = np.min, np.max
    hi = fmax(input_image)
    keep_cval = (
        mode == "constant"
        and not (lo <= cval <= hi)
        and (fmin(output_image) <= cval <= fmax(output_image))
    )
    if keep_cval:
        mask = output_image == cval
    np.clip(output_image, lo, hi, out=output_image)
    if keep_cval:
        output_image[mask] = cval
    return output_image


# --------------------------------------------------------------------------- warp
def _keys_cubic(x, f0, f1, f2, f3):
    """Catmull-Rom / Keys(a=-1/2) in the exact operation order of skimage's
    ``cubic_interpolation`` (identical to ``src/bicubic_interpolation.py:39-41``)."""
    return f1 + 0.5 * x * (
        f2 - f0 + x * (2.0 * f0 - 5.0 * f1 + 4.0 * f2 - f3 + x * (3.0 * (f1 - f2) + f3 - f0))
    )


def project_grid(matrix, rows, cols):
    """Output pixel (row i, col j) -> input coordinates (c, r): skimage
    ``_transform_projective`` applied to x=j, y=i."""
    m = np.asarray(matrix, dtype=np.float64)
    jj = np.arange(cols, dtype=np.float64)[None, :]
    ii = np.arange(rows, dtype=np.float64)[:, None]
    xx = m[0, 0] * jj + m[0, 1] * ii + m[0, 2]
    yy = m[1, 0] * jj + m[1, 1] * ii + m[1, 2]
    zz = m[2, 0] * jj + m[2, 1] * ii + m[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        return xx / zz, yy / zz


def _taps(plane, rr, cc, cval):
    rows, cols = plane.shape
    inside = (rr >= 0) & (rr < rows) & (cc >= 0) & (cc < cols)
    vals = plane[np.clip(rr, 0, rows - 1), np.clip(cc, 0, cols - 1)]
    return np.where(inside, vals, cval)


def warp(image, matrix, order=3, mode="constant", cval=0.0, clip=True):
    """``skimage.transform.warp`` for a 3x3 ``matrix`` mapping output (col,row)
    to input (col,row); orders 1 and 3, ``mode='constant'`` only."""
    if mode != "constant":
        raise NotImplementedError("only mode='constant' is on the reference's path")
    img = np.asarray(image, dtype=np.float64)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    rows, cols, nch = img.shape
    c, r = project_grid(matrix, rows, cols)
    finite = np.isfinite(c) & np.isfinite(r)
    c_safe = np.where(finite, c, -1e9)
    r_safe = np.where(finite, r, -1e9)
    # keep the integer parts inside int64 range whatever the warp does
    c_safe = np.clip(c_safe, -1e9, 1e9)
    r_safe = np.clip(r_safe, -1e9, 1e9)
    r0 = np.floor(r_safe)
    c0 = np.floor(c_safe)
    xr = r_safe - r0
    xc = c_safe - c0
    r0 = r0.astype(np.int64)
    c0 = c0.astype(np.int64)
    out = np.empty_like(img)
    for ch in range(nch):
        plane = img[:, :, ch]
        if order == 3:
            fr = []
            for a in range(4):
                f = [_taps(plane, r0 - 1 + a, c0 - 1 + b, cval) for b in range(4)]
                fr.append(_keys_cubic(xc, *f))
            out[:, :, ch] = _keys_cubic(xr, *fr)
        elif order == 1:
            tl = _taps(plane, r0, c0, cval)
            tr_ = _taps(plane, r0, c0 + 1, cval)
            bl = _taps(plane, r0 + 1, c0, cval)
            br = _taps(plane, r0 + 1, c0 + 1, cval)
            top = (1 - xc) * tl + xc * tr_
            bot = (1 - xc) * bl + xc * br
            out[:, :, ch] = (1 - xr) * top + xr * bot
        else:
            raise NotImplementedError("order must be 1 or 3")
    clip_like_skimage(img, out, mode, cval, clip)
    return out[:, :, 0] if squeeze else out


# ------------------------------------------------------------------------ rescale
def rescale(image, scale, order=3, mode="constant", cval=0, clip=True, anti_aliasing=True):
    """``skimage.transform.rescale(..., channel_axis=2, preserve_range=True)``:
    Gaussian(sigma=(f-1)/2, zero extension) -> scipy cubic-spline ``zoom`` with
    ``grid_mode=True`` and ``mode='grid-constant'`` -> clip to the input range."""
    if mode != "constant":
        raise NotImplementedError("only mode='constant' is on the reference's path")
    img = np.asarray(image, dtype=np.float64)
    in_shape = np.asarray(img.shape, dtype=np.float64)
    sc = np.array([scale, scale, 1.0])
    out_shape = np.maximum(np.round(sc * in_shape), 1)[:2]  # np.round: half-to-even
    out_shape = tuple(int(v) for v in out_shape) + (img.shape[2],)
    factors = np.divide(img.shape, out_shape)
    if anti_aliasing:
        sigma = np.maximum(0, (factors - 1) / 2)
        filtered = ndi.gaussian_filter(img, sigma, cval=cval, mode="grid-constant")
    else:
        filtered = img
    out = ndi.zoom(
        filtered, [1 / f for f in factors], order=order, mode="grid-constant", cval=cval,
        grid_mode=True,
    )
    clip_like_skimage(img, out, mode, cval, clip)
    return out
