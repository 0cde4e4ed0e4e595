#!/usr/bin/env python
"""One iteration of the robust loop on an 8192 x 8192 gray pair (BASELINE config 5's size) with the row-blocked
float64 oracle (``ica_oracle.hessian_b_rowblocked``: no 12.9 GB DIJ): H, b and dp = H^-1 b.

    python oracle/make_golden_8192.py        # writes tests/golden/hb_8192.npz (a few minutes of CPU)

The GPU test regenerates the pair (``synthetic.make_large_gray_pair``, deterministic) and compares
``ica_hessian_b_host`` with this record: that is where fp32 per-lane x-moment sums (x^4 ~ 4.5e15) would show.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ica_oracle as orc  # noqa: E402
from inverse_compositional_algorithm_b200 import synthetic  # noqa: E402

CASES = {
    # name: (seed, H, W, p, robust, lambda, delta)
    "8192_lorentzian": (7, 8192, 8192, [1e-4, -5e-5, 2.3, 8e-5, -1.2e-4, -2.6, 1e-9, -2e-9], orc.LORENTZIAN, 40.0, 10),
}


def main():
    out = {"names": np.array(list(CASES))}
    for name, (seed, H, W, p, rt, lam, delta) in CASES.items():
        I1, I2 = synthetic.make_large_gray_pair(seed, H, W)
        t0 = time.perf_counter()
        Hm, b = orc.hessian_b_rowblocked(I1, I2, np.array(p), orc.HOMOGRAPHY, rt, lam, True, delta, channel_mult=3.0, rows=64)
        dp = np.linalg.solve(Hm, b)
        print(f"{name}: {time.perf_counter() - t0:.0f} s, dp = {dp}", flush=True)
        out[name + "/cfg"] = np.array([seed, H, W, rt, lam, delta], dtype=np.float64)
        out[name + "/p"] = np.array(p)
        out[name + "/H"] = Hm
        out[name + "/b"] = b
        out[name + "/dp"] = dp
        out[name + "/checksum"] = np.array([I1.sum(dtype=np.float64), I2.sum(dtype=np.float64)])
    path = os.path.join(ROOT, "tests", "golden", "hb_8192.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
