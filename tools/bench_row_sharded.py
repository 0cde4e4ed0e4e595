#!/usr/bin/env python
"""BASELINE.json configs[4]: ONE 8192x8192 grayscale pair, homography, row-sharded over the ranks with a
per-iteration NCCL allreduce of the moment sums (H/b), SURVEY.md 8e.

    python tools/bench_row_sharded.py                       # 1 GPU (no collective)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/bench_row_sharded.py      # N GPUs

Prints one JSON line on rank 0: time per registration (CUDA events, max over ranks), iterations, and the
latency of the allreduce (CUDA events around every call).  Not the headline bench (that is bench.py)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--nscales", type=int, default=5)
    ap.add_argument("--robust", default="LORENTZIAN")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--max-lin", type=float, default=0.002, help="size of the linear part of the ground-truth motion")
    ap.add_argument("--emulate", type=int, default=0, help="emulate this many ranks on one GPU (no NCCL)")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer"],
                    help="nccl: host-driven loop with one NCCL allreduce per iteration; peer: exchange inside the "
                         "device-side loop through peer-mapped buffers (NVLink), one graph launch per registration")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.image_optimisation import RobustErrorFunctionType
    from inverse_compositional_algorithm_b200.sharding import register_row_sharded
    from inverse_compositional_algorithm_b200.transformation import TransformType

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t = TransformType.HOMOGRAPHY
    n = args.size
    # the same pair on every rank (same seed); ground truth known
    # ground truth: a few pixels of shift, a linear part and a perspective part that move the far corner by ~10 px each
    lin, persp = 10.0 / n, 10.0 / (n * n)
    p_true = [0.6 * lin, -0.4 * lin, 5.3, 0.5 * lin, -0.7 * lin, -3.7, 0.6 * persp, -0.5 * persp]
    I1, I2, p_gt = synthetic.make_batch_device(1, n, n, 1, t, seed=5, device="cuda", p_gt=[p_true])
    I1, I2 = I1[0].contiguous(), I2[0].contiguous()
    rt = RobustErrorFunctionType[args.robust]
    times, ar, iters_all = [], [], None
    for rep in range(args.reps + 1):
        stats = {"time": True}
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p, err, iters = register_row_sharded(I1, I2, t, nscales=args.nscales, robust_type=rt, delta=10,
                                             emulate_ranks=args.emulate or None, stats=stats, exchange=args.exchange)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rep > 0:            # the first repetition warms NCCL and the plan up
            times.append(float(ms.item()))
            ar.extend(stats["allreduce_ms"])
        iters_all = iters
    if rank == 0:
        from inverse_compositional_algorithm_b200.transformation import project_points
        cx, cy = np.meshgrid(np.linspace(0, n - 1, 9), np.linspace(0, n - 1, 9))
        xa, ya = project_points(cx, cy, p[:8], t)
        xb, yb = project_points(cx, cy, p_gt[0], t)
        epe = float(np.hypot(xa - xb, ya - yb).max())      # on a 9x9 grid of control points
        ar = np.asarray(ar)
        print(json.dumps({
            "workload": f"single {n}x{n} grayscale pair, homography, {args.robust}, {args.nscales} scales, "
                        f"row-sharded over {args.emulate or world} ranks" + (" (emulated on one GPU)" if args.emulate else ""),
            "n_gpus": world, "exchange": args.exchange, "ms_per_registration": float(np.median(times)), "reps": args.reps,
            "iterations": int(stats["iterations"]), "launched_iterations": int(stats["launched_iterations"]), "iters_per_scale(coarse->fine)": [int(v) for v in iters_all[::-1]],
            "allreduce_ms_median": float(np.median(ar)), "allreduce_ms_mean": float(ar.mean()),
            "allreduce_ms_p95": float(np.percentile(ar, 95)), "allreduce_bytes": 105 * 8,
            "exchange_us_mean(peer mode: publish -> all ranks seen)": stats.get("exchange_us_mean"),
            "epe_vs_ground_truth_px": epe}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
