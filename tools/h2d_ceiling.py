#!/usr/bin/env python
"""Host->device copy ceiling of the box: plain ``cudaMemcpyAsync`` from pinned host memory on N GPUs at once (one
process per GPU, one copy per call -- no batched-copy API).  This is the bound of every end-to-end number that feeds
8-bit image pairs from host memory (bench.py's ``e2e``): at 1024 x 1024 x 3 a pair is 6.3 MB, so R registrations/s per
GPU need 6.3 R MB/s of link and of host-memory read bandwidth.

    python tools/h2d_ceiling.py                                            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        tools/h2d_ceiling.py                                               # eight GPUs concurrently

Prints one JSON line on rank 0: per-rank and aggregate GB/s for a few copy sizes."""
import json
import os

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1 and hasattr(os, "sched_setaffinity"):
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
    out = {"n_gpus": world, "sizes": {}}
    for mb in (48, 201, 805):          # 8, 32, 128 pairs of 2 x 1024 x 1024 x 3 bytes
        n = mb * 1000 * 1000
        src = torch.empty(n, dtype=torch.uint8).pin_memory()
        src.fill_(rank + 1)
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        reps = max(3, 2000 // mb)
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)     # one cudaMemcpyAsync each
        e1.record()
        torch.cuda.synchronize()
        gbs = torch.tensor([n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9], dtype=torch.float64, device="cuda")
        allg = [torch.zeros_like(gbs) for _ in range(world)]
        if world > 1:
            dist.all_gather(allg, gbs)
        else:
            allg = [gbs]
        vals = [float(v.item()) for v in allg]
        out["sizes"][f"{mb}MB"] = {"per_rank_GBs": [round(v, 2) for v in vals], "aggregate_GBs": round(sum(vals), 1),
                                   "min_rank_GBs": round(min(vals), 2)}
        del src, dst
    if rank == 0:
        out["pairs_per_s_ceiling_per_gpu(6.29 MB per 1024^2 RGB u8 pair)"] = round(out["sizes"]["201MB"]["min_rank_GBs"] * 1e9 / 6291456)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
