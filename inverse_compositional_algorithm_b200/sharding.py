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
