"""Multi-GPU partitioning of a batch of independent registrations.

The hot path shards by image pair with NO data-path collective: rank r of W owns a contiguous
block of pairs, registers them on its own GPU with its own plan, and only the small per-pair
results (8 parameters, error, iteration counts) are gathered.  The only collectives are that
gather and the timing barrier; they go through ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_pairs: int, rank: int, world: int):
    """Contiguous block ``[lo, hi)`` of rank ``rank``; block sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_pairs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def register_sharded(I1, I2, transform_types, register_fn, *, group=None):
    """Runs ``register_fn(I1[lo:hi], I2[lo:hi], types[lo:hi]) -> (p [n,8], err [n], iters [n,S])`` on this
    rank's block and all-gathers the results so that every rank returns the full arrays.
    ``I1``/``I2`` may be the full batch (every rank slices its block) -- no image crosses ranks."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(I1)
    lo, hi = shard_range(n, rank, world)
    types = list(transform_types) if isinstance(transform_types, (list, tuple, np.ndarray)) else [transform_types] * n
    p, err, iters = register_fn(I1[lo:hi], I2[lo:hi], types[lo:hi])
    if world == 1:
        return p, err, iters
    nscales = iters.shape[1]
    packed = np.zeros((hi - lo, 8 + 1 + nscales))
    packed[:, :8], packed[:, 8], packed[:, 9:] = p, err, iters
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    sizes = [shard_range(n, r, world) for r in range(world)]
    maxn = max(b - a for a, b in sizes)
    buf = torch.zeros((maxn, packed.shape[1]), dtype=torch.float64, device=dev)
    buf[:hi - lo] = torch.from_numpy(packed).to(dev)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    full = np.concatenate([o.cpu().numpy()[:b - a] for o, (a, b) in zip(out, sizes)])
    return full[:, :8], full[:, 8], full[:, 9:].astype(np.int32)


# ---------------------------------------------------------------------------------------------
# Row-sharded registration of ONE large pair (BASELINE.json configs[4]).
# The reference sums H (ica.py:95-99) and b (ica.py:101, 240-246) over all pixels; that sum is the only
# global step of an iteration.  Each rank gathers the moment sums of its band of rows on its own GPU, the
# ranks add them with one allreduce (NCCL: 105 doubles per pair), and every rank performs the identical
# solve / update on the identical sums.  Both full images live on every rank.
# ---------------------------------------------------------------------------------------------
_ROW_PLANS = {}   # plans of register_row_sharded, reused between calls (device buffers, pyramid operators)


def row_band(tiles_y: int, rank: int, nranks: int):
    """Band ``[ty0, ty1)`` of tile rows that ``rank`` owns (same arithmetic as the kernels)."""
    import ctypes as C
    from . import _native
    a, b = C.c_int32(), C.c_int32()
    _native.check(_native.lib().ica_row_band(int(tiles_y), int(rank), int(nranks), C.byref(a), C.byref(b)))
    return a.value, b.value


def register_row_sharded(I1, I2, transform_type, *, nscales=5, nu=0.5, robust_type=0, robust_loop=None,
                         lambda_=0.0, tol=1e-3, max_iter=30, delta=5, nanifoutside=True,
                         gray_as_rgb=True, p0=None, group=None, emulate_ranks=None, stats=None, poll_every=4,
                         exchange="nccl"):
    """Registers one image pair with the pixels of every iteration split by rows over the ranks of
    ``group`` (``torch.distributed``, NCCL).  ``I1``/``I2``: float32 CUDA tensors ``[H, W, C]`` (the full
    images, on every rank).  Returns ``(p [8], err, iters [nscales])``, identical on all ranks.

    ``exchange="nccl"`` (default) drives the loop from the host: per iteration K2 on the band, one NCCL allreduce of the
    105 moment sums, K3.  ``exchange="peer"`` moves the exchange into the device-side loop: the ranks map one another's
    exchange buffers (CUDA IPC over NVLink, set up once per plan through ``group``) and the solve kernel adds the moment
    sums through them -- one CUDA-graph launch per registration, no collective call and no host round trip per
    iteration; ``stats`` then receives ``exchange_us_mean`` (publish -> every rank's sums seen).

    ``emulate_ranks=W`` runs W bands one after the other on the current GPU and adds their moments
    locally -- the same kernels and band arithmetic without a process group (used by the tests).
    ``stats`` (a dict) receives ``iterations`` and, when timing was requested through
    ``stats={"time": True}``, the CUDA-event time of every allreduce in ms (``allreduce_ms``)."""
    import torch
    import torch.distributed as dist
    from . import _native
    from .image_optimisation import RobustErrorFunctionType

    for I in (I1, I2):
        if not (isinstance(I, torch.Tensor) and I.is_cuda and I.dtype == torch.float32 and I.dim() == 3):
            raise ValueError("I1 and I2 must be float32 CUDA tensors of the same [H, W, C] shape")
    if I1.shape != I2.shape or I1.device != I2.device:
        raise ValueError("I1 and I2 must be float32 CUDA tensors of the same [H, W, C] shape on one device")
    I1, I2 = I1.contiguous(), I2.contiguous()
    H, W, Cn = (int(v) for v in I1.shape)
    rt = RobustErrorFunctionType(getattr(robust_type, "value", robust_type)).value
    if robust_loop is None:
        robust_loop = rt != 0
    if exchange not in ("nccl", "peer"):
        raise ValueError("exchange must be 'nccl' or 'peer'")
    if exchange == "peer":
        if emulate_ranks:
            raise ValueError("the peer exchange waits on other GPUs inside a kernel: it cannot be emulated on one GPU")
        return _register_row_sharded_peer(I1, I2, transform_type, nscales=nscales, nu=nu, rt=rt, robust_loop=robust_loop,
                                          lambda_=lambda_, tol=tol, max_iter=max_iter, delta=delta, nanifoutside=nanifoutside,
                                          gray_as_rgb=gray_as_rgb, p0=p0, group=group, stats=stats)
    if emulate_ranks:
        world, ranks = int(emulate_ranks), list(range(int(emulate_ranks)))
    else:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        ranks = [dist.get_rank(group) if dist.is_initialized() else 0]
    plans = []
    for r in ranks:
        key = (torch.cuda.current_device(), H, W, Cn, nscales, nu, int(getattr(transform_type, "value", transform_type)),
               rt, bool(robust_loop), lambda_, tol, max_iter, delta, bool(nanifoutside), bool(gray_as_rgb), r, world)
        if key in _ROW_PLANS:
            plans.append(_ROW_PLANS[key])
            continue
        pl = _native.Plan(batch=1, height=H, width=W, channels=Cn, nscales=nscales, nu=nu,
                          transform_type=int(getattr(transform_type, "value", transform_type)), robust_type=rt, robust_loop=bool(robust_loop),
                          lambda_=lambda_, tol=tol, max_iter=max_iter, delta=delta,
                          nanifoutside=nanifoutside, gray_as_rgb=bool(gray_as_rgb) and Cn == 1)
        pl.set_row_shard(r, world)
        if len(_ROW_PLANS) >= 16:
            _ROW_PLANS.pop(next(iter(_ROW_PLANS))).close()
        _ROW_PLANS[key] = pl
        plans.append(pl)
    dev = I1.device
    stride = int(_native.lib().ica_moment_stride())
    p_dev = torch.zeros((1, 8), dtype=torch.float64, device=dev)
    if p0 is not None:
        p0 = np.asarray(p0, dtype=np.float64)
        p_dev[0, :p0.size] = torch.from_numpy(p0).to(dev)
    moments = [torch.zeros((1, stride), dtype=torch.float64, device=dev) for _ in plans]
    stream = torch.cuda.current_stream().cuda_stream
    want_time = bool(stats is not None and stats.get("time"))
    ar_events = []
    for pl in plans:
        pl.shard_begin(I1.data_ptr(), I2.data_ptr(), p_dev.data_ptr(), stream)
    n_active, it = 1, 0
    while n_active > 0 and it < nscales * max_iter:
        for pl, m in zip(plans, moments):
            pl.shard_partial(m.data_ptr(), stream)
        if want_time:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if emulate_ranks:
            total = moments[0]
            for m in moments[1:]:
                total = total + m          # fixed order, like a ring over ranks 0..W-1
        else:
            total = moments[0]
            if world > 1:
                dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        if want_time:
            e1.record()
            ar_events.append((e0, e1))
        # the number of unfinished pairs is read back (a stream synchronisation) only every few iterations: an
        # iteration past convergence has an empty work list and changes nothing
        poll = (it + 1) % poll_every == 0 and it + 1 >= nscales     # a pair needs at least one iteration per scale
        for pl in plans:
            n = pl.shard_solve(total.data_ptr(), stream, poll=poll)
            if n >= 0:
                n_active = n
        it += 1
    out = torch.zeros((1, 8), dtype=torch.float64, device=dev)
    for pl in plans:
        pl.shard_finish(out.data_ptr(), stream)
    torch.cuda.current_stream().synchronize()
    p, err, iters = plans[0].results()
    if stats is not None:
        stats["iterations"] = int(iters.sum())
        stats["launched_iterations"] = it
        stats["world"] = world
        if want_time:
            stats["allreduce_ms"] = [a.elapsed_time(b) for a, b in ar_events]
    return p[0], float(err[0]), iters[0]


def _register_row_sharded_peer(I1, I2, transform_type, *, nscales, nu, rt, robust_loop, lambda_, tol, max_iter, delta,
                               nanifoutside, gray_as_rgb, p0, group, stats):
    """``register_row_sharded(..., exchange="peer")``: plan with a peer-mapped exchange buffer per rank, the whole
    registration as one graph launch (C-ABI ``ica_plan_xchg_create / _connect / ica_plan_run_row_sharded``)."""
    import torch
    import torch.distributed as dist
    from . import _native
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    H, W, Cn = (int(v) for v in I1.shape)
    tt = int(getattr(transform_type, "value", transform_type))
    key = ("peer", torch.cuda.current_device(), H, W, Cn, nscales, nu, tt, rt, bool(robust_loop), lambda_, tol, max_iter, delta,
           bool(nanifoutside), bool(gray_as_rgb), rank, world)
    pl = _ROW_PLANS.get(key)
    if pl is None:
        pl = _native.Plan(batch=1, height=H, width=W, channels=Cn, nscales=nscales, nu=nu, transform_type=tt, robust_type=rt,
                          robust_loop=bool(robust_loop), lambda_=lambda_, tol=tol, max_iter=max_iter, delta=delta,
                          nanifoutside=nanifoutside, gray_as_rgb=bool(gray_as_rgb) and Cn == 1)
        handle = pl.xchg_create(world, rank)
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=group)
            pl.xchg_connect(b"".join(handles))
            dist.barrier(group=group)          # every rank has mapped every buffer before anybody writes
        if len(_ROW_PLANS) >= 16:
            _ROW_PLANS.pop(next(iter(_ROW_PLANS))).close()
        _ROW_PLANS[key] = pl
    dev = I1.device
    p_dev = torch.zeros((1, 8), dtype=torch.float64, device=dev)
    if p0 is not None:
        p0 = np.asarray(p0, dtype=np.float64)
        p_dev[0, :p0.size] = torch.from_numpy(p0).to(dev)
    stream = torch.cuda.current_stream()
    pl.run_row_sharded(I1.data_ptr(), I2.data_ptr(), p_dev.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    p, err, iters = pl.results()
    if stats is not None:
        st = pl.xchg_stats()
        if st["error"]:
            raise RuntimeError("row-sharded registration: a peer did not answer within the exchange time-out")
        stats["iterations"] = int(iters.sum())
        stats["launched_iterations"] = int(st["count"])
        stats["world"] = world
        stats["exchange_us_mean"] = st["mean_us"]
        stats["allreduce_ms"] = [st["mean_us"] * 1e-3] * max(1, int(st["count"]))
    return p[0], float(err[0]), iters[0]
