"""Latency of ONE call of the drop-in driver (what a user of the reference's API sees), 1024x1024 RGB, homography,
Lorentzian, 5 scales: float64 / uint8 numpy inputs, DI and Iw returned."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inverse_compositional_algorithm_b200 import synthetic
from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (
    pyramidal_inverse_compositional_algorithm, register_batch)
from inverse_compositional_algorithm_b200.transformation import TransformType

t = TransformType.HOMOGRAPHY
I1, I2, p_gt = synthetic.make_pair(7, 1024, 1024, 3, t)
for name, a, b in (("float64", np.round(I1).astype(np.float64), np.round(I2).astype(np.float64)),
                   ("uint8", np.round(I1).astype(np.uint8), np.round(I2).astype(np.uint8))):
    ts = []
    for rep in range(6):
        t0 = time.perf_counter()
        p, err, DI, Iw = pyramidal_inverse_compositional_algorithm(a, b, np.zeros(8), t, 5, 0.5, 1e-3, 3, 0.0, True, 10, False)
        ts.append(time.perf_counter() - t0)
    print(f"{name}: full driver call with DI/Iw: first {ts[0]*1e3:.1f} ms, then median {np.median(ts[1:])*1e3:.2f} ms")
    ts = []
    for rep in range(6):
        t0 = time.perf_counter()
        register_batch(a[None], b[None], t, nscales=5, robust_type=3, delta=10)
        ts.append(time.perf_counter() - t0)
    print(f"{name}: register_batch of one pair (parameters only): median {np.median(ts[1:])*1e3:.2f} ms")
