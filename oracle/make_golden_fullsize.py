#!/usr/bin/env python
"""Golden runs at the FULL sizes of BASELINE.json's configs, produced by the CPU oracle (test infrastructure:
``oracle/ica_oracle.py``, the function-by-function restatement of the reference pinned by ``make_golden.py``).

    python oracle/make_golden_fullsize.py            # writes tests/golden/fullsize_runs.npz  (about 3 minutes)

The GPU parity tests regenerate the same seeded inputs (``synthetic.make_pair``: numpy + scipy, deterministic)
and compare the CUDA path's final parameters, per-scale iteration counts and per-iteration ``|dp|`` with these
records in milliseconds, without running the oracle on the GPU box.

Runs (8-bit quantised images, TOL 1e-3, nu 0.5, delta 10, lambda schedule):
  c3_sim, c3_aff   640x480 gray (as its RGB replication), SIMILARITY / AFFINITY, QUADRATIC, 5 scales
  c4_a, c4_b       1024x1024 RGB, HOMOGRAPHY, GERMAN_MCCLURE, 20 % occlusion, 5 scales
  c2_a             1024x1024 RGB, HOMOGRAPHY, LORENTZIAN, 5 scales
  c2_diverge       same, with a ground-truth motion far outside the capture range: 30 iterations at every scale
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ica_oracle as orc  # noqa: E402
from inverse_compositional_algorithm_b200 import synthetic  # noqa: E402
from inverse_compositional_algorithm_b200.transformation import TransformType  # noqa: E402

T = TransformType
# name: (seed, H, W, C, transform, robust, occlusion, make_pair kwargs)
RUNS = {
    "c3_sim": (300, 480, 640, 1, T.SIMILARITY, orc.QUADRATIC, 0.0, {}),
    "c3_aff": (301, 480, 640, 1, T.AFFINITY, orc.QUADRATIC, 0.0, {}),
    "c4_a": (300, 1024, 1024, 3, T.HOMOGRAPHY, orc.GERMAN_MCCLURE, 0.2, {}),
    "c4_b": (301, 1024, 1024, 3, T.HOMOGRAPHY, orc.GERMAN_MCCLURE, 0.2, {}),
    "c2_a": (1, 1024, 1024, 3, T.HOMOGRAPHY, orc.LORENTZIAN, 0.0, {}),
    "c2_diverge": (2, 1024, 1024, 3, T.HOMOGRAPHY, orc.LORENTZIAN, 0.0,
                   dict(margin=192, p_gt=[0.03, -0.02, 170.0, 0.025, -0.03, -150.0, 2e-6, -3e-6])),
}
NSCALES, NU, TOL, DELTA, LAMBDA = 5, 0.5, 1e-3, 10, 0.0


def make_inputs(name):
    seed, H, W, C, t, rt, occ, kw = RUNS[name]
    I1, I2, p_gt = synthetic.make_pair(seed, H, W, C, t, occlusion=occ, **kw)
    return np.round(I1), np.round(I2), p_gt


def main():
    out = {"names": np.array(list(RUNS))}
    for name, (seed, H, W, C, t, rt, occ, kw) in RUNS.items():
        I1, I2, p_gt = make_inputs(name)
        a, b = (np.repeat(x, 3, 2) if C == 1 else x for x in (I1.astype(np.float64), I2.astype(np.float64)))
        trace = []
        t0 = time.perf_counter()
        p, err, _, _ = orc.ica_pyramidal(a, b, np.zeros(t.nparams()), t.value, NSCALES, NU, TOL, rt, LAMBDA, True, DELTA,
                                         trace=trace)
        dt = time.perf_counter() - t0
        n = t.nparams()
        traj = np.zeros((len(trace), 4 + 8))
        for i, (s, it, dpn, pk, lam) in enumerate(trace):
            traj[i, 0], traj[i, 1], traj[i, 2] = s, it, dpn
            traj[i, 3] = np.nan if lam is None else lam
            traj[i, 4:4 + n] = pk
        iters = np.array([sum(1 for r in trace if r[0] == s) for s in range(NSCALES)], dtype=np.int32)
        out[name + "/p"] = p
        out[name + "/err"] = np.float64(err)
        out[name + "/iters"] = iters
        out[name + "/traj"] = traj
        out[name + "/p_gt"] = p_gt
        out[name + "/checksum"] = np.array([I1.sum(dtype=np.float64), I2.sum(dtype=np.float64)])   # the inputs were regenerated identically
        print(f"{name}: {len(trace)} iterations {iters.tolist()} |dp|={err:.3e} in {dt:.1f} s, EPE vs ground truth "
              f"{orc.end_point_error(p, p_gt, t.value, W, H)[1]:.3e} px", flush=True)
    path = os.path.join(ROOT, "tests", "golden", "fullsize_runs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
