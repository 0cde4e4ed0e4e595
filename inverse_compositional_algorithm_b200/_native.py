"""ctypes binding of ``libica_b200.so`` (C-ABI in ``include/ica_b200.h``).

This is the only door from Python to the arithmetic: there is NO CPU fallback.  If the shared
library is missing it is built in-tree with nvcc (``build.py``); if that is impossible, or a
compute entry point is called without a visible GPU, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import build as _build

MAX_PARAMS = 8
MAX_SCALES = 12
TRAJ_STRIDE = 12

FLAG_RECORD_TRAJECTORY = 1
FLAG_WRITE_DI_IW = 2
FLAG_HOST_LOOP = 8
FLAG_IPOL_PYRAMID = 16   # zoom.zoom_out levels instead of skimage rescale
FLAG_IPOL_WARP = 32      # warp domain of bicubic_interpolation_image instead of skimage.transform.warp's

DTYPE_F32, DTYPE_U8, DTYPE_F64 = 0, 1, 2
DTYPE_RGB_TO_LUMA = 0x10     # modifier: RGB host images, one-channel plan registers their luminance
DTYPE_OUT_F64 = 0x20         # modifier: DI / Iw come back as float64 (widened on the device)

ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_ALLOC = -1, -2, -3, -4


class Config(C.Structure):
    """Mirror of ``struct ica_config``."""
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("channels", C.c_int32),
        ("gray_as_rgb", C.c_int32), ("nscales", C.c_int32), ("nu", C.c_double),
        ("transform_type", C.c_int32), ("robust_type", C.c_int32), ("robust_loop", C.c_int32),
        ("lambda_", C.c_double), ("tol", C.c_double), ("max_iter", C.c_int32), ("delta", C.c_int32),
        ("nanifoutside", C.c_int32), ("flags", C.c_uint32), ("blocks_per_pair", C.c_int32),
    ]


# every symbol include/ica_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_PD = C.POINTER(C.c_double)
_PI = C.POINTER(C.c_int32)
_PF = C.POINTER(C.c_float)
SIGNATURES = {
    "ica_last_error": (C.c_char_p, []),
    "ica_version": (C.c_int, []),
    "ica_device_count": (C.c_int, []),
    "ica_set_device": (C.c_int, [C.c_int]),
    "ica_get_device": (C.c_int, [_PI]),
    "ica_get_constants": (C.c_int, [_PD]),
    "ica_plan_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "ica_plan_destroy": (C.c_int, [_P]),
    "ica_plan_set_transform_types": (C.c_int, [_P, _PI, C.c_int32]),
    "ica_plan_level_shapes": (C.c_int, [_P, _PI, _PI]),
    "ica_plan_device_bytes": (C.c_size_t, [_P]),
    "ica_plan_run_device": (C.c_int, [_P, _P, _P, _P, _P]),
    "ica_moment_stride": (C.c_int, []),
    "ica_row_band": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _PI, _PI]),
    "ica_plan_set_row_shard": (C.c_int, [_P, C.c_int32, C.c_int32]),
    "ica_plan_shard_begin": (C.c_int, [_P, _P, _P, _P, _P]),
    "ica_plan_shard_partial": (C.c_int, [_P, _P, _P]),
    "ica_plan_shard_solve": (C.c_int, [_P, _P, _PI, _P]),
    "ica_plan_shard_finish": (C.c_int, [_P, _P, _P]),
    "ica_plan_xchg_create": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "ica_plan_xchg_connect": (C.c_int, [_P, _P]),
    "ica_plan_run_row_sharded": (C.c_int, [_P, _P, _P, _P, _P]),
    "ica_plan_xchg_stats": (C.c_int, [_P, _PD, C.POINTER(C.c_int64), _PI]),
    "ica_plan_run_host": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P, _P]),
    "ica_plan_last_host_run_ms": (C.c_int, [_P, _PF]),
    "ica_plan_debug_timeline": (C.c_int, [_P, _P, C.c_int32]),
    "ica_plan_get_results": (C.c_int, [_P, _P, _P, _P]),
    "ica_plan_get_trajectory": (C.c_int, [_P, _P, _P]),
    "ica_plan_get_di_iw_device": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "ica_plan_get_level_device": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P), _PI]),
    "ica_plan_last_launch_count": (C.c_int64, [_P]),
    "ica_plan_enable_timing": (C.c_int, [_P, C.c_int32]),
    "ica_plan_get_timing": (C.c_int, [_P, _PF, _PI, _PF, _PI]),
    "ica_rhop_host": (C.c_int, [_P, C.c_int64, C.c_double, C.c_int32, _P]),
    "ica_robust_error_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, _P]),
    "ica_steepest_descent_host": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "ica_dij_reduce_host": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "ica_transform_image_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "ica_apply_operators_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _PI, _P, C.c_int32, C.c_int32,
                                           _PI, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "ica_warp_ipol_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "ica_warp_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "ica_rescale_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_double, _P, _PI, _PI]),
    "ica_resample_operator": (C.c_int, [C.c_int32, C.c_int32, _PI, _PI, _P, C.c_int32, _PI]),
    "ica_zoom_size": (C.c_int, [C.c_int32, C.c_int32, C.c_double, _PI, _PI]),
    "ica_generate_pairs_device": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _PI, _P, _PI, C.c_int32,
                                            C.c_uint64, C.c_int32, C.c_int32, C.c_double, C.c_int32, _P]),
    "ica_zoom_out_operator": (C.c_int, [C.c_int32, C.c_double, _PI, _PI, _PI, _P, C.c_int32]),
    "ica_zoom_out_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_double, _P, _PI, _PI]),
    "ica_gradient_host": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "ica_hessian_b_host": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P,
                                     C.c_int32, C.c_double, C.c_int32, C.c_int32, _P, _P]),
    "ica_nparams": (C.c_int, [C.c_int32]),
    "ica_params2matrix": (C.c_int, [_P, C.c_int32, _P]),
    "ica_update_transform": (C.c_int, [_P, _P, C.c_int32]),
    "ica_zoom_in_parameters": (C.c_int, [_P, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, _P]),
    "ica_inverse_hessian": (C.c_int, [_P, C.c_int32, _P]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Loads (building first if needed) the native library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        override = os.environ.get("ICA_LIB_PATH")      # tuning hook: a library variant built elsewhere (tools/build_variant.sh)
        if override:
            if not os.path.exists(override):
                raise RuntimeError(f"ICA_LIB_PATH={override} does not exist")
            path = override
        # rebuild when missing, or when a source is newer than the library and a compiler is at hand (on a box
        # without nvcc the shipped library is used as it is)
        if not override and (not os.path.exists(path) or (_build.is_stale() and _build.have_nvcc())):
            try:
                path = _build.build()
            except Exception as exc:  # noqa: BLE001
                raise RuntimeError(
                    "libica_b200.so is missing and could not be built; the B200 path has no CPU "
                    f"fallback ({exc})") from exc
        handle = C.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().ica_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc == 0:
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(f"libica_b200 error {rc}: {msg}")


def device_count() -> int:
    return int(lib().ica_device_count())


def set_device(index: int) -> None:
    check(lib().ica_set_device(int(index)))


def current_device() -> int:
    d = C.c_int32(-1)
    check(lib().ica_get_device(C.byref(d)))
    return d.value


def require_gpu() -> None:
    if device_count() < 1:
        raise RuntimeError("no CUDA device visible: inverse_compositional_algorithm_b200 has no "
                           "CPU fallback")


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_P)


# ------------------------------------------------------------------ scalar algebra (host side)
def nparams(ttype: int) -> int:
    n = lib().ica_nparams(int(ttype))
    if n < 0:
        raise ValueError("Unknown transform type")
    return n


def params2matrix(p: np.ndarray, ttype: int) -> np.ndarray:
    pp = np.zeros(MAX_PARAMS)
    pp[:min(len(p), MAX_PARAMS)] = p[:MAX_PARAMS]
    m = np.empty(9)
    check(lib().ica_params2matrix(_ptr(pp), int(ttype), _ptr(m)))
    return m.reshape(3, 3)


def update_transform(p: np.ndarray, dp: np.ndarray, ttype: int) -> np.ndarray:
    n = nparams(ttype)
    pp = np.zeros(MAX_PARAMS)
    dd = np.zeros(MAX_PARAMS)
    pp[:n] = p[:n]
    dd[:n] = dp[:n]
    check(lib().ica_update_transform(_ptr(pp), _ptr(dd), int(ttype)))
    return pp[:n].copy()


def zoom_in_parameters(p: np.ndarray, ttype: int, nx, ny, nxx, nyy) -> np.ndarray:
    n = nparams(ttype)
    pp = np.zeros(MAX_PARAMS)
    pp[:n] = p[:n]
    out = np.zeros(MAX_PARAMS)
    check(lib().ica_zoom_in_parameters(_ptr(pp), int(ttype), float(nx), float(ny), float(nxx),
                                       float(nyy), _ptr(out)))
    return out[:n].copy()


def zoom_size(nx: int, ny: int, factor: float):
    a, b = C.c_int32(), C.c_int32()
    check(lib().ica_zoom_size(int(nx), int(ny), float(factor), C.byref(a), C.byref(b)))
    return a.value, b.value


def resample_operator(n_in: int, n_out: int):
    """Dense (n_out x n_in) matrix of the banded pyramid operator along one axis, plus the uniform range."""
    taps = C.c_int32()
    start = np.zeros(n_out, dtype=np.int32)
    check(lib().ica_resample_operator(int(n_in), int(n_out), C.byref(taps), None, None, 0, None))   # band width first
    w = np.zeros(n_out * taps.value, dtype=np.float32)
    fast = np.zeros(3, dtype=np.int32)
    check(lib().ica_resample_operator(int(n_in), int(n_out), C.byref(taps), start.ctypes.data_as(_PI), _ptr(w), w.size,
                                      fast.ctypes.data_as(_PI)))
    A = np.zeros((n_out, n_in))
    for o in range(n_out):
        A[o, start[o]:start[o] + taps.value] = w[o * taps.value:(o + 1) * taps.value]
    return A, taps.value, fast


def zoom_out_operator(n_in: int, factor: float):
    """``(start [n_out] int32, weights [n_out, taps] float32)`` of the zoom_out operator along one axis (built in C++)."""
    n_out, taps = C.c_int32(), C.c_int32()
    check(lib().ica_zoom_out_operator(int(n_in), float(factor), C.byref(n_out), C.byref(taps), None, None, 0))
    start = np.zeros(n_out.value, dtype=np.int32)
    w = np.zeros(n_out.value * taps.value, dtype=np.float32)
    check(lib().ica_zoom_out_operator(int(n_in), float(factor), C.byref(n_out), C.byref(taps), start.ctypes.data_as(_PI),
                                      _ptr(w), w.size))
    return start, w.reshape(n_out.value, taps.value)


def zoom_out(image, factor: float) -> np.ndarray:
    require_gpu()
    img = _as_image_f32(image)
    nxx, nyy = zoom_size(img.shape[1], img.shape[0], factor)
    out = np.empty((max(nyy, 1), max(nxx, 1), img.shape[2]), dtype=np.float32)
    oh, ow = C.c_int32(), C.c_int32()
    check(lib().ica_zoom_out_host(_ptr(img), img.shape[0], img.shape[1], img.shape[2], float(factor), _ptr(out),
                                  C.byref(oh), C.byref(ow)))
    assert (oh.value, ow.value) == out.shape[:2]
    return out


def generate_pairs_device(I1_ptr: int, I2_ptr: int, batch: int, height: int, width: int, channels: int, ttypes, p_gt,
                          occ_xy=None, occ_side: int = 0, seed: int = 0, pair_offset: int = 0, margin: int = 64,
                          noise_sigma: float = 1.0, quantize: bool = True, stream: int = 0) -> None:
    """Fills the device buffers ``I1``/``I2`` (float32 ``[B][H][W][C]``) with synthetic pairs (csrc/ica_generate.cu)."""
    require_gpu()
    tt = np.ascontiguousarray(ttypes, dtype=np.int32)
    pg = np.zeros((batch, MAX_PARAMS))
    pgi = np.asarray(p_gt, dtype=np.float64).reshape(batch, -1)
    pg[:, :pgi.shape[1]] = pgi
    occ = np.ascontiguousarray(occ_xy, dtype=np.int32) if occ_xy is not None else None
    check(lib().ica_generate_pairs_device(_P(I1_ptr), _P(I2_ptr), int(batch), int(height), int(width), int(channels),
                                          tt.ctypes.data_as(_PI), _ptr(pg), occ.ctypes.data_as(_PI) if occ is not None else None,
                                          int(occ_side), C.c_uint64(int(seed)), int(pair_offset), int(margin), float(noise_sigma),
                                          1 if quantize else 0, _P(stream)))


def inverse_hessian(H: np.ndarray) -> np.ndarray:
    H = np.ascontiguousarray(H, dtype=np.float64)
    n = H.shape[0]
    out = np.empty((n, n))
    check(lib().ica_inverse_hessian(_ptr(H), n, _ptr(out)))
    return out


def constants() -> np.ndarray:
    out = np.empty(5)
    check(lib().ica_get_constants(out.ctypes.data_as(_PD)))
    return out


# ------------------------------------------------------------------ stateless GPU helpers
def _as_image_f32(image) -> np.ndarray:
    img = np.ascontiguousarray(image, dtype=np.float32)
    if img.ndim == 2:
        img = img[:, :, None]
    if img.ndim != 3 or img.shape[2] not in (1, 3):
        raise ValueError("image must be (H, W), (H, W, 1) or (H, W, 3)")
    return img


# ---- helper API on materialised float64 arrays (include/ica_b200.h, last section)
def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def rhop(t2, lambda_: float, robust_type: int) -> np.ndarray:
    require_gpu()
    a = _f64(t2)
    out = np.empty_like(a)
    check(lib().ica_rhop_host(_ptr(a.reshape(-1)), a.size, float(lambda_), int(robust_type), _ptr(out.reshape(-1))))
    return out


def robust_error(DI, lambda_: float, robust_type: int) -> np.ndarray:
    require_gpu()
    d = _f64(DI)
    ny, nx, nz = d.shape
    rho = np.empty((ny, nx))
    check(lib().ica_robust_error_host(_ptr(d), ny, nx, nz, float(lambda_), int(robust_type), _ptr(rho)))
    return rho


def steepest_descent(Ix, Iy, J, nparams: int) -> np.ndarray:
    require_gpu()
    ix, iy, j = _f64(Ix), _f64(Iy), _f64(J)
    ny, nx, nz = ix.shape
    out = np.empty((ny, nx, nz, nparams))
    check(lib().ica_steepest_descent_host(_ptr(ix), _ptr(iy), _ptr(j), ny, nx, nz, int(nparams), _ptr(out)))
    return out


def dij_reduce(DIJ, DI=None, rho=None) -> np.ndarray:
    """H (DI is None) or b (DI given), optionally rho-weighted."""
    require_gpu()
    dij = _f64(DIJ)
    ny, nx, nz, n = dij.shape
    if ny * nx * nz == 0:      # empty image: the sum over no pixels (the reference's einsum returns zeros)
        return np.zeros(n if DI is not None else (n, n))
    di = _f64(DI) if DI is not None else None
    r = _f64(rho) if rho is not None else None
    out = np.empty(n if di is not None else (n, n))
    check(lib().ica_dij_reduce_host(_ptr(dij), _ptr(di) if di is not None else None, _ptr(r) if r is not None else None,
                                    ny, nx, nz, n, _ptr(out)))
    return out


def transform_image(image, inverse_matrix) -> np.ndarray:
    require_gpu()
    img = _f64(image)
    squeeze = img.ndim == 2
    if squeeze:
        img = img[:, :, None]
    m = _f64(inverse_matrix).reshape(9)
    out = np.empty_like(img)
    check(lib().ica_transform_image_host(_ptr(img), img.shape[0], img.shape[1], img.shape[2], _ptr(m), _ptr(out)))
    return out[:, :, 0] if squeeze else out


def apply_operators(image, ystart, yweights, xstart, xweights, clip: bool) -> np.ndarray:
    """``out = A_y . image . A_x^T`` with banded operators given as (start [n_out], weights [n_out, taps])."""
    require_gpu()
    img = _as_image_f32(image)
    ys = np.ascontiguousarray(ystart, dtype=np.int32); yw = np.ascontiguousarray(yweights, dtype=np.float32)
    xs = np.ascontiguousarray(xstart, dtype=np.int32); xw = np.ascontiguousarray(xweights, dtype=np.float32)
    out = np.empty((ys.size, xs.size, img.shape[2]), dtype=np.float32)
    check(lib().ica_apply_operators_host(_ptr(img), img.shape[0], img.shape[1], img.shape[2],
                                         ys.ctypes.data_as(_PI), _ptr(yw), yw.shape[1], ys.size,
                                         xs.ctypes.data_as(_PI), _ptr(xw), xw.shape[1], xs.size, 1 if clip else 0, _ptr(out)))
    return out


def warp_ipol(image, params, nparams: int, nanifoutside: bool, delta: int) -> np.ndarray:
    require_gpu()
    img = _f64(image)
    p = np.zeros(8)
    p[:nparams] = np.asarray(params, dtype=np.float64)[:nparams]
    out = np.empty_like(img)
    check(lib().ica_warp_ipol_host(_ptr(img), img.shape[0], img.shape[1], img.shape[2], _ptr(p), int(nparams),
                                   1 if nanifoutside else 0, int(delta), _ptr(out)))
    return out


def warp(image, matrix) -> np.ndarray:
    require_gpu()
    img = _as_image_f32(image)
    m = np.ascontiguousarray(matrix, dtype=np.float64).reshape(9)
    out = np.empty_like(img)
    check(lib().ica_warp_host(_ptr(img), img.shape[0], img.shape[1], img.shape[2], _ptr(m), _ptr(out)))
    return out


def rescale(image, nu: float) -> np.ndarray:
    require_gpu()
    img = _as_image_f32(image)
    nxx, nyy = zoom_size(img.shape[1], img.shape[0], nu)
    out = np.empty((max(nyy, 1), max(nxx, 1), img.shape[2]), dtype=np.float32)
    oh, ow = C.c_int32(), C.c_int32()
    check(lib().ica_rescale_host(_ptr(img), img.shape[0], img.shape[1], img.shape[2], float(nu),
                                 _ptr(out), C.byref(oh), C.byref(ow)))
    assert (oh.value, ow.value) == out.shape[:2]
    return out


def gradient(image, delta: int, nanifoutside: bool):
    require_gpu()
    img = _as_image_f32(image)
    ix = np.empty_like(img)
    iy = np.empty_like(img)
    check(lib().ica_gradient_host(_ptr(img), img.shape[0], img.shape[1], img.shape[2], int(delta),
                                  1 if nanifoutside else 0, _ptr(ix), _ptr(iy)))
    return ix, iy


def hessian_b(I1, I2, ttype: int, p, robust_type: int, lambda_: float, delta: int,
              nanifoutside: bool, gray_as_rgb: bool = False):
    require_gpu()
    a = _as_image_f32(I1)
    b = _as_image_f32(I2)
    if a.shape != b.shape:
        raise ValueError("I1 and I2 must have the same dimensions")
    n = nparams(ttype)
    pp = np.zeros(MAX_PARAMS)
    pp[:n] = np.asarray(p, dtype=np.float64)[:n]
    H = np.zeros((n, n))
    bv = np.zeros(n)
    check(lib().ica_hessian_b_host(_ptr(a), _ptr(b), a.shape[0], a.shape[1], a.shape[2],
                                   1 if gray_as_rgb else 0, int(ttype), _ptr(pp), int(robust_type),
                                   float(lambda_), int(delta), 1 if nanifoutside else 0, _ptr(H),
                                   _ptr(bv)))
    return H, bv


# ------------------------------------------------------------------ plan
class Plan:
    """Owner of one ``ica_plan`` (pyramids, per-pair state and partial buffers for B pairs)."""

    def __init__(self, *, batch, height, width, channels, nscales, nu, transform_type, robust_type,
                 robust_loop, lambda_, tol, max_iter, delta, nanifoutside, gray_as_rgb=False,
                 record_trajectory=False, write_di_iw=False, blocks_per_pair=0, host_loop=False, ipol_pyramid=False,
                 ipol_warp=False):
        require_gpu()
        flags = (FLAG_RECORD_TRAJECTORY if record_trajectory else 0) | (
            FLAG_WRITE_DI_IW if write_di_iw else 0) | (FLAG_HOST_LOOP if host_loop else 0) | (
            FLAG_IPOL_PYRAMID if ipol_pyramid else 0) | (FLAG_IPOL_WARP if ipol_warp else 0)
        self.cfg = Config(batch=batch, height=height, width=width, channels=channels,
                          gray_as_rgb=1 if gray_as_rgb else 0, nscales=nscales, nu=nu,
                          transform_type=int(transform_type), robust_type=int(robust_type),
                          robust_loop=1 if robust_loop else 0, lambda_=lambda_, tol=tol,
                          max_iter=max_iter, delta=delta, nanifoutside=1 if nanifoutside else 0,
                          flags=flags, blocks_per_pair=blocks_per_pair)
        self._h = _P()
        check(lib().ica_plan_create(C.byref(self.cfg), C.byref(self._h)))
        self.batch, self.height, self.width, self.channels = batch, height, width, channels
        self.nscales, self.max_iter = nscales, max_iter
        self.record_trajectory, self.write_di_iw = record_trajectory, write_di_iw

    def close(self):
        if getattr(self, "_h", None):
            lib().ica_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def handle(self):
        return self._h

    def set_transform_types(self, types):
        t = np.ascontiguousarray(types, dtype=np.int32)
        check(lib().ica_plan_set_transform_types(self._h, t.ctypes.data_as(_PI), t.size))

    def level_shapes(self):
        nx = np.zeros(self.nscales, dtype=np.int32)
        ny = np.zeros(self.nscales, dtype=np.int32)
        check(lib().ica_plan_level_shapes(self._h, nx.ctypes.data_as(_PI), ny.ctypes.data_as(_PI)))
        return nx, ny

    def device_bytes(self) -> int:
        return int(lib().ica_plan_device_bytes(self._h))

    def last_launch_count(self) -> int:
        return int(lib().ica_plan_last_launch_count(self._h))

    def enable_timing(self, enable=True):
        check(lib().ica_plan_enable_timing(self._h, int(enable)))

    def timing(self):
        a, b = C.c_float(), C.c_float()
        na, nb = C.c_int32(), C.c_int32()
        check(lib().ica_plan_get_timing(self._h, C.byref(a), C.byref(na), C.byref(b), C.byref(nb)))
        return {"iterate_ms": a.value, "iterate_launches": na.value, "pyramid_ms": b.value,
                "pyramid_launches": nb.value}

    def run_device(self, I1_ptr: int, I2_ptr: int, p_ptr: int, stream: int = 0):
        """Device pointers (ints): float32 [B][H][W][C] x2, double [B][8]."""
        check(lib().ica_plan_run_device(self._h, _P(I1_ptr), _P(I2_ptr), _P(p_ptr), _P(stream)))

    # ---- row-sharded mode (one large pair over several ranks; include/ica_b200.h)
    def set_row_shard(self, rank: int, nranks: int):
        check(lib().ica_plan_set_row_shard(self._h, int(rank), int(nranks)))

    def shard_begin(self, I1_ptr: int, I2_ptr: int, p_ptr: int, stream: int = 0):
        check(lib().ica_plan_shard_begin(self._h, _P(I1_ptr), _P(I2_ptr), _P(p_ptr), _P(stream)))

    def shard_partial(self, moments_ptr: int, stream: int = 0):
        check(lib().ica_plan_shard_partial(self._h, _P(moments_ptr), _P(stream)))

    def shard_solve(self, moments_ptr: int, stream: int = 0, poll: bool = True) -> int:
        n = C.c_int32(-1)
        check(lib().ica_plan_shard_solve(self._h, _P(moments_ptr), C.byref(n) if poll else None, _P(stream)))
        return n.value

    def shard_finish(self, p_ptr: int, stream: int = 0):
        check(lib().ica_plan_shard_finish(self._h, _P(p_ptr), _P(stream)))

    # ---- row-sharded mode, exchange inside the device loop (peer memory over NVLink)
    def xchg_create(self, world: int, rank: int) -> bytes:
        h = (C.c_ubyte * 64)()
        check(lib().ica_plan_xchg_create(self._h, int(world), int(rank), C.cast(h, _P)))
        return bytes(h)

    def xchg_connect(self, handles: bytes):
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        check(lib().ica_plan_xchg_connect(self._h, C.cast(buf, _P)))

    def run_row_sharded(self, I1_ptr: int, I2_ptr: int, p_ptr: int, stream: int = 0):
        check(lib().ica_plan_run_row_sharded(self._h, _P(I1_ptr), _P(I2_ptr), _P(p_ptr), _P(stream)))

    def xchg_stats(self):
        us, n, err = C.c_double(), C.c_int64(), C.c_int32()
        check(lib().ica_plan_xchg_stats(self._h, C.byref(us), C.byref(n), C.byref(err)))
        return {"mean_us": us.value, "count": n.value, "error": err.value}

    def results(self):
        p = np.zeros((self.batch, MAX_PARAMS))
        err = np.zeros(self.batch)
        iters = np.zeros((self.batch, self.nscales), dtype=np.int32)
        check(lib().ica_plan_get_results(self._h, _ptr(p), _ptr(err), _ptr(iters)))
        return p, err, iters

    def run_host(self, I1: np.ndarray, I2: np.ndarray, p0=None, want_images=False, rgb_to_luma=False, images_f64=False):
        """``I1``/``I2``: arrays [B][H][W][C] of dtype float32, uint8 or float64 (C-contiguous).  With ``rgb_to_luma`` the
        inputs are RGB ``[B][H][W][3]`` and the (one-channel) plan registers their luminance, computed on the device."""
        shape = (self.batch, self.height, self.width, self.channels)
        in_shape = (self.batch, self.height, self.width, 3) if rgb_to_luma else shape
        if rgb_to_luma and self.channels != 1:
            raise ValueError("rgb_to_luma needs a one-channel plan")
        if I1.shape != in_shape or I2.shape != in_shape:
            raise ValueError(f"expected image batches of shape {in_shape}, got {I1.shape} / {I2.shape}")
        if I1.dtype != I2.dtype:
            raise ValueError("I1 and I2 must share a dtype")
        code = {np.dtype(np.float32): DTYPE_F32, np.dtype(np.uint8): DTYPE_U8,
                np.dtype(np.float64): DTYPE_F64}.get(I1.dtype)
        if code is None:
            I1 = I1.astype(np.float32)
            I2 = I2.astype(np.float32)
            code = DTYPE_F32
        I1 = np.ascontiguousarray(I1)
        I2 = np.ascontiguousarray(I2)
        p = np.zeros((self.batch, MAX_PARAMS))
        if p0 is not None:
            p0 = np.asarray(p0, dtype=np.float64).reshape(self.batch, -1)
            p[:, :p0.shape[1]] = p0
        err = np.zeros(self.batch)
        iters = np.zeros((self.batch, self.nscales), dtype=np.int32)
        DI = Iw = None
        di_ptr = iw_ptr = None
        if want_images:
            DI = np.empty(shape, dtype=np.float64 if images_f64 else np.float32)
            Iw = np.empty(shape, dtype=np.float64 if images_f64 else np.float32)
            di_ptr, iw_ptr = _ptr(DI), _ptr(Iw)
            if images_f64:
                code |= DTYPE_OUT_F64
        check(lib().ica_plan_run_host(self._h, _ptr(I1), _ptr(I2), code | (DTYPE_RGB_TO_LUMA if rgb_to_luma else 0),
                                      _ptr(p), _ptr(err), _ptr(iters), di_ptr, iw_ptr))
        return p, err, iters, DI, Iw

    def run_host_ptrs(self, I1_ptr: int, I2_ptr: int, dtype_code: int, p: np.ndarray,
                      err: np.ndarray, iters: np.ndarray, DI_ptr: int = 0, Iw_ptr: int = 0):
        """Raw host pointers (e.g. pinned torch tensors) -- used by bench.py's e2e legs."""
        check(lib().ica_plan_run_host(self._h, _P(I1_ptr), _P(I2_ptr), int(dtype_code), _ptr(p),
                                      _ptr(err), _ptr(iters), _P(DI_ptr) if DI_ptr else None, _P(Iw_ptr) if Iw_ptr else None))

    def last_host_run_ms(self) -> float:
        ms = C.c_float()
        check(lib().ica_plan_last_host_run_ms(self._h, C.byref(ms)))
        return ms.value

    def debug_timeline(self, enable=True, fetch=False):
        grid = lib().ica_plan_debug_timeline(self._h, None, 1 if enable else 0)
        if not fetch:
            return None
        out = np.zeros((grid, 16), dtype=np.int64)
        lib().ica_plan_debug_timeline(self._h, _ptr(out), 1 if enable else 0)
        return out

    def trajectory(self):
        """List (per pair) of arrays [count][12]: scale, iter, |dp|, lambda, p[8]."""
        cap = self.nscales * self.max_iter
        traj = np.zeros((self.batch, cap, TRAJ_STRIDE))
        cnt = np.zeros(self.batch, dtype=np.int32)
        check(lib().ica_plan_get_trajectory(self._h, _ptr(traj), _ptr(cnt)))
        return [traj[b, :cnt[b]].copy() for b in range(self.batch)]

    def level_device(self, which: int, pair: int, scale: int):
        ptr, pitch = _P(), C.c_int32()
        check(lib().ica_plan_get_level_device(self._h, which, pair, scale, C.byref(ptr), C.byref(pitch)))
        return ptr.value, pitch.value
