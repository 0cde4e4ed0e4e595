#!/usr/bin/env python
"""Golden outputs of the reference's IPOL-style warp ``bicubic_interpolation_image`` (src/bicubic_interpolation.py:
121-152, numba) -> tests/golden/ipol_warp.npz.  Runs the UNMODIFIED reference in the build container (it cannot
travel to the GPU box).  Test infrastructure only."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import reference_loader  # noqa: E402

CASES = [  # (params, nanifoutside, delta)
    ([1.3, -0.7], True, 2),
    ([2.25, 1.5, 0.04], True, 0),
    ([-1.0, 0.5, 0.02, -0.03], False, 3),
    ([0.4, -0.6, 0.01, 0.02, -0.015, 0.03], True, 1),
    ([0.01, 0.02, 1.5, -0.01, 0.015, -2.0, 1e-4, -2e-4], True, 2),
    ([-3.6, -2.2], False, -5),          # negative delta: negative coordinates reach the sign-dependent stencil
]


def main():
    warnings.simplefilter("ignore")
    bi = reference_loader.load()["bi"]
    rng = np.random.default_rng(20240826)
    img = rng.uniform(0, 255, (23, 31, 3))
    out = {"image": img}
    for i, (p, nan_out, delta) in enumerate(CASES):
        out[f"params_{i}"] = np.asarray(p, dtype=np.float64)
        out[f"flags_{i}"] = np.asarray([1 if nan_out else 0, delta])
        out[f"out_{i}"] = bi.bicubic_interpolation_image(img, np.asarray(p, dtype=np.float64), len(p), nan_out, delta)
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "ipol_warp.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
