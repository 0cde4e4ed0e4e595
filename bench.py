#!/usr/bin/env python
"""Benchmark of the inverse compositional hot path (BASELINE.json metric):
registrations/sec, 1024x1024 RGB, HOMOGRAPHY (8 parameters), LORENTZIAN, 5-scale pyramid.

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --steps K --warmup W    # CPU oracle port on the host cores

One "step" = one pass of the hot path (both pyramids + the whole coarse-to-fine robust loop)
over one batch of B synthetic image pairs per GPU.  `value` is measured with the inputs already
resident in HBM (CUDA events, max over ranks); `e2e` goes through the host-buffer C-ABI entry
(`ica_plan_run_host`, what the Python drop-in calls) with pinned host inputs, host->device and
device->host copies inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "c2": dict(name="1024x1024 RGB pairs, homography (8 params), Lorentzian robust error, 5-scale pyramid",
               H=1024, W=1024, C=3, transform="HOMOGRAPHY", robust="LORENTZIAN", nscales=5, occlusion=0.0),
    # configs[3]: Geman-McClure with 20% occlusion
    "c4": dict(name="1024x1024 RGB pairs, homography, Geman-McClure, 20% occlusion, 5-scale pyramid",
               H=1024, W=1024, C=3, transform="HOMOGRAPHY", robust="GERMAN_MCCLURE", nscales=5, occlusion=0.2),
    # configs[2]: 640x480 gray, similarity/affinity mix, quadratic
    "c3": dict(name="640x480 grayscale pairs, similarity/affinity mix, quadratic error, 5-scale pyramid",
               H=480, W=640, C=1, transform="MIX", robust="QUADRATIC", nscales=5, occlusion=0.0),
}
NU, TOL, DELTA, LAMBDA = 0.5, 1e-3, 10, 0.0
DATA_SEED = 20260              # the one synthetic data set of all arms: pair k = pair k of set DATA_SEED (synthetic.py)
METRIC = "registrations/sec (1024^2 homography, 5-scale robust)"
UNIT = "pairs/s"


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(iters, nx, ny, C):
    """per pixel-iteration the fused kernel must read I1 once and I2 once."""
    px = (nx.astype(np.int64) * ny.astype(np.int64))[None, :]
    return float((iters.astype(np.int64) * px).sum() * 2 * C * 4), float((iters.astype(np.int64) * px).sum())


# --------------------------------------------------------------------------- CPU oracle legs
def _oracle_one(args):
    """Worker: one registration with the CPU oracle port (numpy/scipy restatement of the reference)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    pair, wl = args[:2]
    from oracle import ica_oracle as orc
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.transformation import TransformType
    from inverse_compositional_algorithm_b200.image_optimisation import RobustErrorFunctionType
    t = pair_type(wl, pair)
    if len(args) > 2:          # images handed over by the caller (the GPU arm's own pair, for the parity re-check)
        I1, I2 = args[2], args[3]
    else:                      # pair `pair` of the data set every arm uses, regenerated on the CPU (numpy mirror of the
        I1, I2, _ = synthetic.make_pair_hash(DATA_SEED, pair, wl["H"], wl["W"], wl["C"], t, occlusion=wl["occlusion"])   # device generator)
    if wl["C"] == 1:
        I1, I2 = np.repeat(I1, 3, 2), np.repeat(I2, 3, 2)
    t0 = time.perf_counter()
    trace = []
    p, _, _, _ = orc.ica_pyramidal(I1, I2, np.zeros(t.nparams()), t.value, wl["nscales"], NU, TOL,
                                   RobustErrorFunctionType[wl["robust"]].value, LAMBDA, True, DELTA, trace=trace)
    return time.perf_counter() - t0, len(trace), p


def pair_type(wl, pair):
    """Transform type of pair `pair` of the data set (config 3 alternates similarity / affinity)."""
    from inverse_compositional_algorithm_b200.transformation import TransformType
    if wl["transform"] == "MIX":
        return TransformType.SIMILARITY if pair % 2 == 0 else TransformType.AFFINITY
    return TransformType[wl["transform"]]


def run_reference(args, wl):
    """`--impl reference`: the reference is Python/numpy and does not travel to the GPU box, so this
    arm times the oracle port (oracle/ica_oracle.py, function-by-function restatement pinned to the
    reference's own outputs) on the host cores, one independent registration per process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    workers = max(1, min(cores, 64))
    try:        # one registration of the oracle holds ~2.5 GB of float64 temporaries at 1024^2 x 3
        import psutil
        workers = max(1, min(workers, int(psutil.virtual_memory().available // (3 << 30))))
    except Exception:  # noqa: BLE001
        workers = min(workers, 16)
    ctx = mp.get_context("spawn")
    budget_s = 240.0
    with ctx.Pool(workers) as pool:
        small = dict(wl, H=128, W=128)
        for _ in range(max(1, min(args.warmup, 2))):       # warm the pool / imports on a tiny sample
            pool.map(_oracle_one, [(i, small) for i in range(workers)])
        t0 = time.perf_counter()
        first = pool.map(_oracle_one, [(i, wl) for i in range(workers)])          # pairs 0.. of the GPU arm's rank-0 batch
        t_step = time.perf_counter() - t0
        steps = max(1, min(args.steps, int(budget_s // max(t_step, 1e-3))))
        total = t_step
        for s in range(1, steps):
            t0 = time.perf_counter()
            pool.map(_oracle_one, [(s * workers + i, wl) for i in range(workers)])
            total += time.perf_counter() - t0
    value = steps * workers / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "pairs_per_step": workers, "nu": NU, "TOL": TOL, "delta": DELTA,
                   "data_set": f"synthetic set {DATA_SEED}, pairs 0..{steps * workers - 1}: the first pairs of the GPU arm's batch, "
                               "regenerated on the CPU by the numpy mirror of the device generator"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": f"{workers} registrations per step, one per process "
                                   f"(single-pair latency {np.mean([r[0] for r in first]):.1f} s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU oracle port of the reference's numpy/scipy path (the Python reference cannot travel to "
                "the GPU box); steps capped so the run ends within minutes",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- CUDA path
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from inverse_compositional_algorithm_b200 import _native, synthetic
    from inverse_compositional_algorithm_b200.transformation import TransformType, end_point_error
    from inverse_compositional_algorithm_b200.image_optimisation import RobustErrorFunctionType

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    _native.set_device(local)
    dev = torch.device("cuda", local)
    # one node, several ranks: give every rank its own slice of the host cores (its e2e leg runs several host threads;
    # without this the ranks' threads migrate over all cores and fight for them)
    affinity = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        mine = cores[local * per:(local + 1) * per] or cores
        try:
            os.sched_setaffinity(0, mine)
            affinity = len(mine)
        except OSError:
            pass

    B, H, W, C, ns = args.batch, wl["H"], wl["W"], wl["C"], wl["nscales"]
    types = [pair_type(wl, rank * B + i) for i in range(B)]
    robust = RobustErrorFunctionType[wl["robust"]]
    # rank r registers pairs [r*B, (r+1)*B) of the data set, generated in place by the library's generator kernels
    # (8-bit values, what the reference's notebooks read from disk, stored as float32)
    I1, I2, p_gt = synthetic.make_batch_device(B, H, W, C, types, seed=DATA_SEED, pair_offset=rank * B, device=dev,
                                               occlusion=wl["occlusion"], quantize=True)
    plan = _native.Plan(batch=B, height=H, width=W, channels=C, nscales=ns, nu=NU, transform_type=types[0].value,
                        robust_type=robust.value, robust_loop=robust != RobustErrorFunctionType.QUADRATIC,
                        lambda_=LAMBDA, tol=TOL, max_iter=30, delta=DELTA, nanifoutside=True, gray_as_rgb=(C == 1), blocks_per_pair=args.chunks)
    plan.set_transform_types([t.value for t in types])
    nx, ny = plan.level_shapes()
    p_dev = torch.zeros((B, 8), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        p_dev.zero_()
        plan.run_device(I1.data_ptr(), I2.data_ptr(), p_dev.data_ptr(), stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput.  The batch is registered as `--streams` independent sub-batches, each with its own
    # plan on its own CUDA stream: the latency-bound phases of one sub-batch (coarse levels, n x n solves) overlap the
    # bandwidth-bound phases of the other.  Everything is enqueued asynchronously (the iteration loop runs on the
    # device inside a CUDA-graph while node); timing is CUDA events on the main stream around all sub-batches.
    S = max(1, min(args.streams, B))
    sb = [(i * B // S, (i + 1) * B // S) for i in range(S)]
    subs = []
    for lo_, hi_ in sb:
        sp = _native.Plan(batch=hi_ - lo_, height=H, width=W, channels=C, nscales=ns, nu=NU,
                          transform_type=types[0].value, robust_type=robust.value,
                          robust_loop=robust != RobustErrorFunctionType.QUADRATIC, lambda_=LAMBDA, tol=TOL, max_iter=30,
                          delta=DELTA, nanifoutside=True, gray_as_rgb=(C == 1), blocks_per_pair=args.chunks)
        sp.set_transform_types([t.value for t in types[lo_:hi_]])
        subs.append(dict(plan=sp, stream=torch.cuda.Stream(device=dev), i1=I1[lo_:hi_], i2=I2[lo_:hi_], p=p_dev[lo_:hi_]))

    def run_steps(nsteps):
        """Enqueues `nsteps` passes over the batch.  Each sub-batch advances through its steps on its own stream
        (in-order per stream, no join between steps), so a straggling pair of one sub-batch overlaps the other
        sub-batches' next step instead of idling the GPU; all streams are joined at the end."""
        fork = torch.cuda.Event()
        fork.record(stream)
        for sub in subs:
            sub["stream"].wait_event(fork)
        for _ in range(nsteps):
            for sub in subs:
                with torch.cuda.stream(sub["stream"]):
                    sub["p"].zero_()
                sub["plan"].run_device(sub["i1"].data_ptr(), sub["i2"].data_ptr(), sub["p"].data_ptr(),
                                       sub["stream"].cuda_stream)
        for sub in subs:
            join = torch.cuda.Event()
            join.record(sub["stream"])
            stream.wait_event(join)

    run_steps(args.warmup)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    run_steps(args.steps)
    ev1.record(stream)
    barrier()
    launches = sum(sub["plan"].last_launch_count() for sub in subs) * args.steps   # identical inputs every step
    elapsed_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    p_streams = p_dev.cpu().numpy().copy()
    # ---- kernel-level numbers for the roofline: one instrumented single-stream step over the whole batch
    # (device-side %globaltimer span of every iterate launch, CUDA events around the pyramid launches)
    plan.enable_timing(1)
    step_device()
    torch.cuda.synchronize()
    tm = plan.timing()
    p_res, err_res, iters = plan.results()
    ab, pi = algorithmic_bytes(iters, nx, ny, C)
    assert np.array_equal(p_res, p_streams), "sub-batched and whole-batch results differ"   # batch-invariant
    # cross-check of the device-side kernel timer: one more step with the host-driven loop and a CUDA-event pair
    # around every iterate launch
    plan.enable_timing(2)
    step_device()
    torch.cuda.synchronize()
    tm_events = plan.timing()
    plan.enable_timing(False)
    per_rank_ms = [elapsed_ms / args.steps]
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(v.item()) / args.steps for v in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry (pinned host memory).  Two plans of B/2 pairs are
    # driven from two host threads so that the host->device copy of one half overlaps the kernels of the
    # other (each ica_plan_run_host call copies its inputs in, runs, copies its results out and synchronises).
    import threading
    host_dtype = torch.uint8 if args.input_dtype == "u8" else torch.float32
    code = _native.DTYPE_U8 if args.input_dtype == "u8" else _native.DTYPE_F32
    nhalf = max(1, min(args.e2e_plans if world == 1 else min(args.e2e_plans, 4), B))
    bounds = [(i * B // nhalf, (i + 1) * B // nhalf) for i in range(nhalf)]
    halves = []
    for lo_, hi_ in bounds:
        nb = hi_ - lo_
        hp = _native.Plan(batch=nb, height=H, width=W, channels=C, nscales=ns, nu=NU, transform_type=types[0].value,
                          robust_type=robust.value, robust_loop=robust != RobustErrorFunctionType.QUADRATIC,
                          lambda_=LAMBDA, tol=TOL, max_iter=30, delta=DELTA, nanifoutside=True, gray_as_rgb=(C == 1), blocks_per_pair=args.chunks)
        hp.set_transform_types([t.value for t in types[lo_:hi_]])
        h1 = torch.empty((nb, H, W, C), dtype=host_dtype).pin_memory()
        h2 = torch.empty((nb, H, W, C), dtype=host_dtype).pin_memory()
        h1.copy_(I1[lo_:hi_].to(host_dtype)); h2.copy_(I2[lo_:hi_].to(host_dtype))
        halves.append(dict(plan=hp, h1=h1, h2=h2, p=np.zeros((nb, 8)), err=np.zeros(nb),
                           it=np.zeros((nb, ns), dtype=np.int32)))
    torch.cuda.synchronize()
    e2e_steps = max(1, 2 * args.steps)     # rounds are short (tens of ms): twice the steps amortise the pipeline fill and drain

    def e2e_worker(hv, nsteps):
        _native.set_device(local)
        for _ in range(nsteps):
            hv["p"][:] = 0
            hv["plan"].run_host_ptrs(hv["h1"].data_ptr(), hv["h2"].data_ptr(), code, hv["p"], hv["err"], hv["it"])

    def e2e_run(nsteps):
        ths = [threading.Thread(target=e2e_worker, args=(hv, nsteps)) for hv in halves]
        for t_ in ths:
            t_.start()
        for t_ in ths:
            t_.join()

    e2e_ms = float("inf")
    if not args.no_e2e:
        e2e_run(max(1, min(args.warmup, 2)))
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        # the e2e leg must reproduce the device-resident results
        p_e2e = np.concatenate([hv["p"] for hv in halves])
        assert np.allclose(p_e2e, p_res, rtol=0, atol=1e-9), "e2e and device-resident results differ"
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    esz = 1 if args.input_dtype == "u8" else 4
    h2d = 2 * B * H * W * C * esz + B * 8 * 8
    d2h = B * 8 * 8 + B * 8 + B * ns * 4
    for hv in halves:
        hv["plan"].close()

    # ---- the reference's full return tuple (p, error, DI, Iw: ica.py:261, 374) end to end: as above plus the two
    # residual / warped images per pair copied device->host (float32, pinned).  A smaller batch (the pinned result
    # buffers take 2 x 12.6 MB per pair); same plans-on-threads scheme.
    full_value, full_pairs, full_d2h = None, 0, 0
    if not args.no_e2e and not args.no_full_tuple:
        nbf = max(1, min(B, args.full_tuple_pairs))
        nplans = max(1, min(4, nbf))
        fb = [(i * nbf // nplans, (i + 1) * nbf // nplans) for i in range(nplans)]
        fulls = []
        for lo_, hi_ in fb:
            nb = hi_ - lo_
            fp = _native.Plan(batch=nb, height=H, width=W, channels=C, nscales=ns, nu=NU, transform_type=types[0].value,
                              robust_type=robust.value, robust_loop=robust != RobustErrorFunctionType.QUADRATIC,
                              lambda_=LAMBDA, tol=TOL, max_iter=30, delta=DELTA, nanifoutside=True, gray_as_rgb=(C == 1),
                              blocks_per_pair=args.chunks, write_di_iw=True)
            fp.set_transform_types([t.value for t in types[lo_:hi_]])
            h1 = torch.empty((nb, H, W, C), dtype=host_dtype).pin_memory()
            h2 = torch.empty((nb, H, W, C), dtype=host_dtype).pin_memory()
            h1.copy_(I1[lo_:hi_].to(host_dtype)); h2.copy_(I2[lo_:hi_].to(host_dtype))
            fulls.append(dict(plan=fp, h1=h1, h2=h2, p=np.zeros((nb, 8)), err=np.zeros(nb), it=np.zeros((nb, ns), dtype=np.int32),
                              di=torch.empty((nb, H, W, C), dtype=torch.float32).pin_memory(),
                              iw=torch.empty((nb, H, W, C), dtype=torch.float32).pin_memory()))
        torch.cuda.synchronize()

        def full_worker(hv, nsteps):
            _native.set_device(local)
            for _ in range(nsteps):
                hv["p"][:] = 0
                hv["plan"].run_host_ptrs(hv["h1"].data_ptr(), hv["h2"].data_ptr(), code, hv["p"], hv["err"], hv["it"],
                                         hv["di"].data_ptr(), hv["iw"].data_ptr())

        def full_run(nsteps):
            ths = [threading.Thread(target=full_worker, args=(hv, nsteps)) for hv in fulls]
            for t_ in ths:
                t_.start()
            for t_ in ths:
                t_.join()

        full_run(1)
        barrier()
        t0 = time.perf_counter()
        full_run(e2e_steps)
        torch.cuda.synchronize()
        full_ms = (time.perf_counter() - t0) * 1e3
        assert np.allclose(np.concatenate([hv["p"] for hv in fulls]), p_res[:nbf], rtol=0, atol=1e-9)
        if world > 1:
            t = torch.tensor([full_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            full_ms = float(t.item())
        full_value = world * nbf * e2e_steps / (full_ms * 1e-3)
        full_pairs, full_d2h = nbf, 2 * nbf * H * W * C * 4 + nbf * (8 * 8 + 8 + ns * 4)
        for hv in fulls:
            hv["plan"].close()
        del fulls

    # ---- accuracy vs ground truth (informative) on this rank's batch
    epe = [end_point_error(p_res[i, :types[i].nparams()], p_gt[i, :types[i].nparams()], types[i], W, H)[1]
           for i in range(min(B, 8))]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    achieved = ab / (tm["iterate_ms"] * 1e-3) / 1e9 if tm["iterate_ms"] > 0 else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "iterate_dram_bytes.json")) as f:
            prof = json.load(f)
            # ncu capture of a 32-pair step: DRAM bytes of the kernel relative to its algorithmic bytes, applied to the
            # per-launch algorithmic bytes of this run (same workload, different batch size)
            traffic = prof["traffic_over_algorithmic"] * ab / max(1, tm["iterate_launches"])
    except Exception:  # noqa: BLE001
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "ms_per_step_per_rank": per_rank_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 images, f64 parameters and reductions", "data": "synthetic",
        "config": {"workload": wl["name"], "pairs_per_step_per_gpu": B, "sub_batches_per_gpu": S, "image_values": "8-bit quantised, float32 in HBM" if args.input_dtype == "u8" else "float32", "nu": NU, "TOL": TOL, "delta": DELTA,
                   "lambda": "schedule 80*0.9^k floored at 5", "l2_policy": "inputs larger than L2 "
                   f"({2 * B * H * W * C * 4 / 2**20:.0f} MiB per step vs 126 MiB)",
                   "iters_per_scale_mean(coarse->fine)": [round(float(v), 2) for v in iters.mean(0)[::-1]]},
        "pixel_iterations_per_s": world * pi * 1.0 / (elapsed_ms / args.steps * 1e-3),
        "epe_vs_ground_truth_px_max": float(np.max(epe)),
        "gpu_launches": int(launches),
        "host_cores_per_rank": affinity,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "entry": f"ica_plan_run_host, pinned {args.input_dtype} host buffers, {nhalf} sub-batch plans "
                f"on {nhalf} host threads (copies take turns on the link and overlap the other plans' kernels), wall clock between device syncs"},
        "e2e_full_tuple": {"value": full_value, "unit": UNIT, "pairs_per_step_per_gpu": full_pairs,
                           "d2h_bytes_per_step": full_d2h,
                           "what": "as e2e, plus the reference's DI and Iw (float32, pinned) copied device->host for every pair"},
        "roofline": {"bound": "hbm", "kernel": "ica_iterate_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "frac_of_nominal_8000_GBs": (achieved / 8000.0) if achieved else None, "traffic": traffic,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_step": ab, "kernel_ms_per_step": tm["iterate_ms"],
                     "launches_per_step": tm["iterate_launches"],
                     "pyramid_ms_per_step": tm["pyramid_ms"],
                     "timer": "one instrumented single-stream step over the whole batch: device %globaltimer span of every iterate launch, accumulated on the device (graph loop)",
                     "kernel_ms_per_step_cuda_events_host_loop": tm_events["iterate_ms"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        # pair 0 of this run's batch through the CPU oracle, on the very images the GPU registered (copied back):
        # the timed CPU baseline and, with it, parity at the workload's full size
        tt = types[0]
        t_cpu, n_it, p_cpu = _oracle_one((0, wl, I1[0].cpu().numpy(), I2[0].cpu().numpy()))
        epe_o = end_point_error(p_res[0, :tt.nparams()], p_cpu, tt, W, H)
        line["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"1 registration of the workload (pair 0 of the batch, {n_it} iterations, {t_cpu:.1f} s), "
                                          "numpy/scipy oracle port, single process",
                                "gpu_vs_oracle_epe_px_mean_max": [epe_o[0], epe_o[1]],
                                "gpu_iterations_same_pair": int(iters[0].sum())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=512, help="image pairs per step per GPU")
    ap.add_argument("--input-dtype", default="u8", choices=["u8", "f32"],
                    help="dtype of the host images on the e2e leg (values are identical on the device-resident leg)")
    ap.add_argument("--chunks", type=int, default=0, help="partial-sum slots per pair (0 = library default: 256, or 1024 for very large images)")
    ap.add_argument("--streams", type=int, default=4, help="independent sub-batches (plan + CUDA stream each) per GPU")
    ap.add_argument("--e2e-plans", type=int, default=8, help="sub-batch plans (one host thread each) on the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    ap.add_argument("--no-full-tuple", action="store_true", help="skip the e2e leg that also returns DI and Iw")
    ap.add_argument("--full-tuple-pairs", type=int, default=64, help="pairs per step per GPU of the e2e_full_tuple leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
