"""Synthetic image pairs with known ground-truth motion.

Two generators of the same kind of pair -- a smooth random texture, its centre crop as I2, and
I1(x) = texture(x'(x; p_gt)) + noise, optionally with a square occlusion:
* :func:`make_pair` (numpy/scipy, cubic-spline resampling): the seeded inputs of the parity tests and golden files;
* :func:`make_batch_device` -- the library's own CUDA generator (``csrc/ica_generate.cu``, counter-based randomness) for
  benchmark-sized batches, with :func:`make_pair_hash`, its numpy mirror, for the CPU legs.  This is input synthesis, not part of the registration path; the reference's
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


def make_large_gray_pair(seed: int, height: int, width: int, shift=(3, -2), noise_sigma: float = 1.0):
    """Cheap deterministic pair for very large single images (BASELINE config 5, 8192 x 8192 gray): an 8-bit smooth
    random texture ``I2`` and ``I1 = I2`` moved by an integer ``shift`` (rows, columns) plus noise, so that
    ``I1(x, y) ~ I2(x - shift[1], y - shift[0])``.  Returns float32 ``(H, W, 1)`` images holding 8-bit values."""
    rng = np.random.default_rng(SEED0 + int(seed))
    tex = rng.standard_normal((height, width), dtype=np.float32)
    tex = ndi.gaussian_filter(tex, sigma=2.0, mode="wrap")
    lo, hi = float(tex.min()), float(tex.max())
    i2 = np.round((tex - lo) * (255.0 / (hi - lo))).astype(np.float32)
    i1 = np.roll(i2, shift, axis=(0, 1)) + noise_sigma * rng.standard_normal((height, width), dtype=np.float32)
    i1 = np.clip(np.round(i1), 0.0, 255.0).astype(np.float32)
    return i1[:, :, None], i2[:, :, None]


# ------------------------------------------------------------------ counter-based generator (device kernel + numpy mirror)
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _gen_key(seed, pair, stream):
    inner = _splitmix64(np.uint64((int(pair) << 8) | int(stream)))
    return _splitmix64(np.uint64(int(seed)) ^ inner)


def _hash_bits(key, idx):
    with np.errstate(over="ignore"):
        return _splitmix64(key + np.asarray(idx, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95))


def _gen_normal(key, idx):
    h = _hash_bits(key, idx)
    u1 = ((h >> np.uint64(40)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
    u2 = ((h >> np.uint64(16)) & np.uint64(0xFFFFFF)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(2.0 * np.pi) * u2)).astype(np.float32)


def _gen_uniform(key, idx):
    return (_hash_bits(key, idx) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def pair_ground_truth(seed: int, pair: int, height: int, width: int, transform_type, *, max_shift=8.0, max_lin=0.02,
                      occlusion=0.0):
    """Ground-truth motion and occlusion corner of pair ``pair`` of the set ``seed`` (host-side numpy stream keyed by
    (seed, pair): every pair can be regenerated alone).  Returns ``(p_gt, (occ_x, occ_y, side))``."""
    rng = np.random.default_rng([SEED0, int(seed), int(pair)])
    p = random_motion(rng, TransformType(transform_type), height, width, max_shift, max_lin)
    side = 0
    ox = oy = 0
    if occlusion > 0:
        side = min(int(round(np.sqrt(occlusion * height * width))), height, width)
        oy = int(rng.integers(0, height - side + 1))
        ox = int(rng.integers(0, width - side + 1))
    return p, (ox, oy, side)


def make_pair_hash(seed: int, pair: int, height: int, width: int, channels: int, transform_type, *, max_shift=8.0,
                   max_lin=0.02, noise_sigma=1.0, occlusion=0.0, margin=64, quantize=True, p_gt=None):
    """numpy mirror of the device generator (``csrc/ica_generate.cu``): the same counter-based noise, blur,
    Catmull-Rom resampling, additive noise and occlusion, so that CPU legs (the oracle, ``bench.py --impl reference``)
    see the images the GPU legs generate in place (equal up to float32 rounding: a handful of grey levels in a
    million 8-bit values).  Returns ``(I1, I2, p_gt)`` float32 ``(H, W, C)``."""
    t = TransformType(transform_type)
    pg, (ox, oy, side) = pair_ground_truth(seed, pair, height, width, t, max_shift=max_shift, max_lin=max_lin,
                                           occlusion=occlusion)
    if p_gt is not None:
        pg = np.asarray(p_gt, dtype=np.float64)
    Ht, Wt, C = height + 2 * margin, width + 2 * margin, channels
    key = _gen_key(seed, pair, 0)
    idx = np.arange(Ht * Wt * C, dtype=np.uint64)
    tex = _gen_normal(key, idx).reshape(Ht, Wt, C)
    k = np.exp(-0.5 * (np.arange(-8, 9, dtype=np.float64) / 2.0) ** 2)
    k = (k / k.sum()).astype(np.float32)
    for axis in (1, 0):          # horizontal pass first, like the kernel; periodic
        acc = np.zeros_like(tex)
        for j in range(17):
            acc = acc + k[j] * np.roll(tex, 8 - j, axis=axis)
        tex = acc
    lo, hi = np.float32(tex.min()), np.float32(tex.max())
    scale = np.float32(255.0) / (hi - lo)
    i2 = (tex[margin:margin + height, margin:margin + width] - lo) * scale
    from .transformation import params2matrix
    m = params2matrix(pg, t)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float64)
    zz = m[2, 0] * xx + m[2, 1] * yy + m[2, 2]
    xs = (m[0, 0] * xx + m[0, 1] * yy + m[0, 2]) / zz + margin
    ys = (m[1, 0] * xx + m[1, 1] * yy + m[1, 2]) / zz + margin
    fx, fy = np.floor(xs), np.floor(ys)
    cx, cy = fx.astype(np.int64), fy.astype(np.int64)
    tx, ty = (xs - fx).astype(np.float32), (ys - fy).astype(np.float32)

    def keys(tt):
        t2 = tt * tt
        return [tt * (tt * (np.float32(-0.5) * tt + 1) - np.float32(0.5)), t2 * (np.float32(1.5) * tt - np.float32(2.5)) + 1,
                tt * (tt * (np.float32(-1.5) * tt + 2) + np.float32(0.5)), t2 * (np.float32(0.5) * tt - np.float32(0.5))]
    wx, wy = keys(tx), keys(ty)
    a = np.zeros((height, width, C), dtype=np.float32)
    for q in range(4):
        rows = np.clip(cy - 1 + q, 0, Ht - 1)
        h = np.zeros((height, width, C), dtype=np.float32)
        for j in range(4):
            cols = np.clip(cx - 1 + j, 0, Wt - 1)
            h = h + wx[j][..., None] * tex[rows, cols]
        a = a + wy[q][..., None] * h
    i1 = (a - lo) * scale
    pidx = np.arange(height * width * C, dtype=np.uint64).reshape(height, width, C)
    if noise_sigma > 0:
        i1 = i1 + np.float32(noise_sigma) * _gen_normal(_gen_key(seed, pair, 1), pidx)
    if side > 0:
        occ = np.float32(255.0) * _gen_uniform(_gen_key(seed, pair, 2), pidx)
        i1[oy:oy + side, ox:ox + side] = occ[oy:oy + side, ox:ox + side]
    i1 = np.clip(i1, 0.0, 255.0)
    if quantize:
        i1, i2 = np.rint(i1), np.rint(i2)
    return i1.astype(np.float32), np.ascontiguousarray(i2, dtype=np.float32), pg


def make_batch_device(batch: int, height: int, width: int, channels: int, transform_types, *, seed: int = 0,
                      pair_offset: int = 0, device="cuda", max_shift=8.0, max_lin=0.02, noise_sigma=1.0, occlusion=0.0,
                      margin=64, quantize=True, p_gt=None):
    """Benchmark-sized batches generated on the GPU by the library's own kernels (``ica_generate_pairs_device``); torch is
    only the container of the device buffers.  Pair ``i`` of the result is pair ``pair_offset + i`` of the set ``seed``
    (the numpy mirror :func:`make_pair_hash` regenerates any of them on the CPU).
    Returns float32 CUDA tensors ``I1, I2`` ``[B, H, W, C]`` and ``p_gt [B, 8]`` (numpy)."""
    import torch
    from . import _native
    if not isinstance(transform_types, (list, tuple)):
        transform_types = [transform_types] * batch
    types = [TransformType(t) for t in transform_types]
    p_all = np.zeros((batch, 8))
    occ_xy = np.zeros((batch, 2), dtype=np.int32)
    side = 0
    for i, t in enumerate(types):
        p, (ox, oy, side) = pair_ground_truth(seed, pair_offset + i, height, width, t, max_shift=max_shift,
                                              max_lin=max_lin, occlusion=occlusion)
        if p_gt is not None:           # explicit ground truth (rows of up to 8 parameters) instead of the drawn one
            p = np.asarray(p_gt, dtype=np.float64).reshape(batch, -1)[i][:t.nparams()]
        p_all[i, :len(p)] = p
        occ_xy[i] = (ox, oy)
    dev = torch.device(device)
    I1 = torch.empty((batch, height, width, channels), device=dev, dtype=torch.float32)
    I2 = torch.empty_like(I1)
    with torch.cuda.device(dev):
        _native.generate_pairs_device(I1.data_ptr(), I2.data_ptr(), batch, height, width, channels,
                                      [t.value for t in types], p_all, occ_xy if side > 0 else None, side, seed=seed,
                                      pair_offset=pair_offset, margin=margin, noise_sigma=noise_sigma, quantize=quantize,
                                      stream=torch.cuda.current_stream(dev).cuda_stream)
    return I1, I2, p_all
