"""Drivers: host-side mirror of the reference's ``src/inverse_compositional_algorithm.py``.

Same three entry points, same argument names, same return tuple ``(p, error, DI, Iw)`` and
the same ``ValueError``s.  All arithmetic runs in ``libica_b200.so`` (sm_100a CUDA): one
``ica_plan_run_host`` call per registration builds both pyramids, iterates every scale on the
device (fused warp / residual / rho' / b / H kernel with the solve-and-compose epilogue) and
returns the parameters; nothing is computed on the CPU and there is no fallback.

Beyond the reference: :func:`register_batch` runs B independent pairs in one call (the batched
shape of the reference's Keras twin, ``src/keras-tf/tf_inverse_compositional_algorithm.py:467-583``,
but with per-pair convergence).
"""
from __future__ import annotations

import threading

import numpy as np

from . import _native
from . import constants as cts
from .image_optimisation import RobustErrorFunctionType, _as_robust
from .transformation import TransformType, _as_type

_PLAN_CACHE: dict = {}
_PLAN_CACHE_MAX = 8
# cached plans are shared by every caller of this module: calls are serialised (a plan runs one registration at a time;
# for concurrent streams of work create `_native.Plan` objects per thread, as bench.py does)
_LOCK = threading.RLock()


def _get_plan(**key):
    # plans, their streams and graphs belong to one CUDA device: the calling thread's current device is part of the key
    key.setdefault("device", _native.current_device())
    k = tuple(sorted(key.items()))
    with _LOCK:
        return _get_plan_locked(k, key)


def _get_plan_locked(k, key):
    plan = _PLAN_CACHE.get(k)
    if plan is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.pop(next(iter(_PLAN_CACHE))).close()
        args = dict(key)
        args.pop("device", None)      # part of the cache key only: plans live on the device that is current
        plan = _native.Plan(**args)
        _PLAN_CACHE[k] = plan
    return plan


def clear_plan_cache():
    while _PLAN_CACHE:
        _PLAN_CACHE.popitem()[1].close()


def _check_rgb(I1, I2):
    # ica.py:48-49, 300-301
    if len(I1.shape) != 3 or len(I2.shape) != 3 or I1.shape[2] != 3 or I2.shape[2] != 3:
        raise ValueError("I1 and I2 must be RGB images with channels in the last dimension")


def _check_common(I1, I2, TOL):
    # ica.py:55-60, 172-177, 304-311
    if I1.shape != I2.shape:
        raise ValueError("I1 and I2 must have the same dimensions")
    if TOL >= 0.01:
        raise ValueError("TOL must be positive and very small (less than 0.01)")


def _as_batch(I):
    I = np.asarray(I)
    if I.dtype not in (np.float32, np.uint8, np.float64):
        I = I.astype(np.float64)  # ica.py:63-65 casts everything to float64
    return np.ascontiguousarray(I)[None]


def _print_trace(traj, n, quadratic, with_scale):
    """The reference's verbose lines (ica.py:125-129, 253-257, 341-345, 357-358)."""
    last_scale = None
    for row in traj:
        s, it, err, lam = int(row[0]), int(row[1]), row[2], row[3]
        if with_scale and s != last_scale:
            print(f"Scale: {s}")
            print("(L2 norm)" if quadratic else "(Robust error function)")
            last_scale = s
        ptxt = " ".join(str(v) for v in row[4:4 + n])
        if quadratic:
            print(f"Iteration {it}: |Dp|={err}: p=({ptxt})")
        else:
            print(f"|Dp|={err}: p=({ptxt}), lambda_={lam}")


def _run_single(I1, I2, p, transform_type, nscales, nu, TOL, robust_type, robust_loop, lambda_,
                nanifoutside, delta, verbose, ipol_pyramid=False, ipol_warp=False):
    t = _as_type(transform_type)
    r = _as_robust(robust_type)
    n = t.nparams()
    ny, nx, nz = I1.shape
    with _LOCK:      # plan acquisition and use are one critical section (an eviction must not close a plan in use)
        plan = _get_plan(batch=1, height=ny, width=nx, channels=nz, nscales=int(nscales), nu=float(nu),
                         transform_type=t.value, robust_type=r.value, robust_loop=bool(robust_loop),
                         lambda_=float(lambda_), tol=float(TOL), max_iter=cts.MAX_ITER, delta=int(delta),
                         nanifoutside=(nanifoutside is True), gray_as_rgb=False,
                         record_trajectory=bool(verbose), write_di_iw=True, ipol_pyramid=bool(ipol_pyramid),
                         ipol_warp=bool(ipol_warp))
        p0 = np.zeros(_native.MAX_PARAMS)
        p0[:n] = np.asarray(p, dtype=np.float64)[:n]
        # DI / Iw come back as float64 like the reference's: widened on the device, copied straight into the result arrays
        pout, err, iters, DI, Iw = plan.run_host(_as_batch(I1), _as_batch(I2), p0[None], want_images=True, images_f64=True)
        if verbose:
            _print_trace(plan.trajectory()[0], n, quadratic=not robust_loop, with_scale=nscales > 1)
    return pout[0, :n].copy(), float(err[0]), DI[0], Iw[0]


def inverse_compositional_algorithm(I1, I2, p, transform_type, TOL, nanifoutside, delta, verbose, *, ipol_warp=False):
    """Quadratic (L2) inverse compositional algorithm, one scale.
    Drop-in for ``src/inverse_compositional_algorithm.py:17-133``; ``p`` is updated in place when
    it is a float64 array (the reference mutates it) and also returned.
    ``ipol_warp=True`` (keyword only, beyond the reference's signature) switches the loop's warp to the IPOL-style
    domain of ``bicubic_interpolation_image``: the run then follows the IPOL C++ logs of the reference's
    ``docs/Algortihm Report.md``."""
    _check_rgb(I1, I2)
    _check_common(I1, I2, TOL)
    pout, err, DI, Iw = _run_single(I1, I2, p, transform_type, 1, 0.5, TOL,
                                    RobustErrorFunctionType.QUADRATIC, False, 0.0, nanifoutside,
                                    delta, verbose, ipol_warp=ipol_warp)
    return _writeback(p, pout), err, DI, Iw


def robust_inverse_compositional_algorithm(I1, I2, p, transform_type, TOL, robust_type, lambda_,
                                           nanifoutside, delta, verbose, *, ipol_warp=False):
    """Robust inverse compositional algorithm, one scale.
    Drop-in for ``src/inverse_compositional_algorithm.py:135-261`` (rho' and the weighted Hessian
    are re-evaluated every iteration, also for QUADRATIC)."""
    _check_common(I1, I2, TOL)
    if len(I1.shape) != 3 or I1.shape[2] not in (1, 3):
        raise ValueError("I1 and I2 must be (H, W, 3) or (H, W, 1) images")
    pout, err, DI, Iw = _run_single(I1, I2, p, transform_type, 1, 0.5, TOL, robust_type, True,
                                    lambda_, nanifoutside, delta, verbose, ipol_warp=ipol_warp)
    return _writeback(p, pout), err, DI, Iw


def pyramidal_inverse_compositional_algorithm(I1, I2, p, transform_type, nscales, nu, TOL,
                                              robust_type, lambda_, nanifoutside, delta, verbose, *, ipol_pyramid=False,
                                              ipol_warp=False):
    """Coarse-to-fine driver.  Drop-in for ``src/inverse_compositional_algorithm.py:264-374``:
    skimage-``rescale`` pyramid, QUADRATIC -> quadratic loop, anything else -> robust loop,
    ``zoom_in_parameters`` between scales; like the reference the input ``p`` is copied and only
    used when ``nscales == 1`` (ica.py:327, 372).
    Keyword-only options beyond the reference's signature (the IPOL-faithful variants the reference
    carries as unused helpers): ``ipol_pyramid`` builds the levels with ``zoom.zoom_out`` instead of skimage ``rescale``,
    ``ipol_warp`` uses the warp domain of ``bicubic_interpolation_image``."""
    _check_rgb(I1, I2)
    _check_common(I1, I2, TOL)
    r = _as_robust(robust_type)
    robust_loop = r != RobustErrorFunctionType.QUADRATIC
    return _run_single(I1, I2, np.copy(p), transform_type, nscales, nu, TOL, r, robust_loop, lambda_,
                       nanifoutside, delta, verbose, ipol_pyramid=ipol_pyramid, ipol_warp=ipol_warp)


def _writeback(p, pout):
    if isinstance(p, np.ndarray) and p.dtype == np.float64 and p.shape == pout.shape:
        p[...] = pout
        return p
    return pout


def register_batch(I1, I2, transform_type, nscales=1, nu=0.5, TOL=1e-3,
                   robust_type=RobustErrorFunctionType.QUADRATIC, lambda_=0.0, nanifoutside=True,
                   delta=10, p0=None, gray_as_rgb=True, return_images=False, plan=None, luminance=False,
                   ipol_pyramid=False, ipol_warp=False):
    """B independent registrations in one call: ``I1``/``I2`` are ``[B, H, W, C]`` (C = 1 or 3,
    float32 / uint8 / float64); ``transform_type`` is one type or a length-B sequence (mixed
    batches).  Returns ``(p [B, 8] zero-padded, error [B], iters [B, nscales])`` and, with
    ``return_images``, ``DI`` and ``Iw``.  A gray batch behaves as the reference would on the
    image replicated to three channels unless ``gray_as_rgb=False``."""
    I1 = np.asarray(I1)
    I2 = np.asarray(I2)
    if I1.ndim != 4 or I1.shape[3] not in (1, 3):
        raise ValueError("I1 and I2 must be [B, H, W, C] with C in (1, 3)")
    _check_common(I1, I2, TOL)
    B, ny, nx, nz = I1.shape
    if luminance:      # RGB in, registered on the luminance (converted on the device from the uploaded RGB bytes)
        if nz != 3:
            raise ValueError("luminance=True needs RGB inputs")
        nz = 1
    types = ([_as_type(transform_type)] * B if not isinstance(transform_type, (list, tuple, np.ndarray))
             else [_as_type(t) for t in transform_type])
    if len(types) != B:
        raise ValueError("one transform type per pair is required")
    r = _as_robust(robust_type)
    with _LOCK:
        if plan is None:
            plan = _get_plan(batch=B, height=ny, width=nx, channels=nz, nscales=int(nscales),
                             nu=float(nu), transform_type=types[0].value, robust_type=r.value,
                             robust_loop=r != RobustErrorFunctionType.QUADRATIC, lambda_=float(lambda_),
                             tol=float(TOL), max_iter=cts.MAX_ITER, delta=int(delta),
                             nanifoutside=(nanifoutside is True), gray_as_rgb=bool(gray_as_rgb) and nz == 1,
                             record_trajectory=False, write_di_iw=bool(return_images), ipol_pyramid=bool(ipol_pyramid),
                             ipol_warp=bool(ipol_warp))
        plan.set_transform_types([t.value for t in types])
        pout, err, iters, DI, Iw = plan.run_host(I1, I2, p0, want_images=return_images, rgb_to_luma=bool(luminance))
    if return_images:
        return pout, err, iters, DI, Iw
    return pout, err, iters


def register_batch_device(I1, I2, transform_type, nscales=1, nu=0.5, TOL=1e-3,
                          robust_type=RobustErrorFunctionType.QUADRATIC, lambda_=0.0, nanifoutside=True,
                          delta=10, p0=None, gray_as_rgb=True):
    """``register_batch`` for images that already live on the GPU: ``I1``/``I2`` are float32 CUDA tensors
    ``[B, H, W, C]`` (torch is only the container: the library reads the device pointers in place, on torch's
    current stream, with no host round trip until the small result arrays are fetched).
    Returns ``(p [B, 8] float64 CUDA tensor, error [B], iters [B, nscales])``."""
    import torch
    for I in (I1, I2):
        if not (isinstance(I, torch.Tensor) and I.is_cuda and I.dtype == torch.float32 and I.dim() == 4):
            raise ValueError("I1 and I2 must be float32 CUDA tensors [B, H, W, C]")
    if I2.device != I1.device:
        raise ValueError("I1 and I2 must live on the same CUDA device")
    if I1.shape[3] not in (1, 3):
        raise ValueError("I1 and I2 must be [B, H, W, C] with C in (1, 3)")
    _check_common(I1, I2, TOL)
    I1, I2 = I1.contiguous(), I2.contiguous()
    B, ny, nx, nz = (int(v) for v in I1.shape)
    types = ([_as_type(transform_type)] * B if not isinstance(transform_type, (list, tuple, np.ndarray))
             else [_as_type(t) for t in transform_type])
    if len(types) != B:
        raise ValueError("one transform type per pair is required")
    r = _as_robust(robust_type)
    with _LOCK, torch.cuda.device(I1.device):
        plan = _get_plan(batch=B, height=ny, width=nx, channels=nz, nscales=int(nscales), nu=float(nu),
                         transform_type=types[0].value, robust_type=r.value,
                         robust_loop=r != RobustErrorFunctionType.QUADRATIC, lambda_=float(lambda_), tol=float(TOL),
                         max_iter=cts.MAX_ITER, delta=int(delta), nanifoutside=(nanifoutside is True),
                         gray_as_rgb=bool(gray_as_rgb) and nz == 1, record_trajectory=False, write_di_iw=False,
                         device=int(I1.device.index or 0))
        plan.set_transform_types([t.value for t in types])
        p = torch.zeros((B, 8), dtype=torch.float64, device=I1.device)
        if p0 is not None:
            p0 = np.asarray(p0, dtype=np.float64).reshape(B, -1)
            p[:, :p0.shape[1]] = torch.from_numpy(p0).to(I1.device)
        stream = torch.cuda.current_stream(I1.device)
        plan.run_device(I1.data_ptr(), I2.data_ptr(), p.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        _, err, iters = plan.results()
    return p, err, iters


class PyramidalInverseCompositional:
    """Batched, layer-shaped front end with the call shape of the reference's Keras layer
    (``tf_inverse_compositional_algorithm.py:467-583``): constructed with the algorithm's options, called on
    ``[I1, I2]`` of shape ``[B, H, W, C]``, returns ``(p [B, 8] zero-padded, error [B], DI, Iw)``.  Unlike the
    reference's layer, whose stopping rule looks at the whole batch (``tf_ica.py:225-232`` -- its authors call it with
    B = 1 for that reason), every pair follows its own coarse-to-fine schedule and stops on its own ``|dp|``.
    No TensorFlow involved: inputs are numpy arrays (uint8 / float32 / float64), the work is the CUDA path."""

    def __init__(self, transform_type, nscales=3, nu=0.5, TOL=1e-3,
                 robust_type=RobustErrorFunctionType.QUADRATIC, lambda_=0.0, nanifoutside=True, delta=10,
                 verbose=False, return_images=True):
        self.transform_type = transform_type
        self.nscales, self.nu, self.TOL = int(nscales), float(nu), float(TOL)
        self.robust_type, self.lambda_ = robust_type, float(lambda_)
        self.nanifoutside, self.delta, self.verbose = nanifoutside, int(delta), bool(verbose)
        self.return_images = bool(return_images)
        self.iterations = None      # [B, nscales] of the last call

    def __call__(self, inputs):
        I1, I2 = inputs
        res = register_batch(I1, I2, self.transform_type, nscales=self.nscales, nu=self.nu, TOL=self.TOL,
                             robust_type=self.robust_type, lambda_=self.lambda_, nanifoutside=self.nanifoutside,
                             delta=self.delta, gray_as_rgb=True, return_images=self.return_images)
        self.iterations = res[2]
        if self.verbose:
            for b in range(len(res[0])):
                print(f"pair {b}: |Dp|={res[1][b]:.6f}: p=({' '.join(f'{v:.6f}' for v in res[0][b])}), "
                      f"iterations per scale (fine->coarse) {res[2][b].tolist()}")
        if self.return_images:
            return res[0], res[1], res[3].astype(np.float64), res[4].astype(np.float64)
        return res[0], res[1]
