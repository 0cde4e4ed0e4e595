"""Accounting of the march kernel over one batched registration (needs a library built with
``ICA_NVCC_EXTRA=-DICA_MARCH_STATS python -m inverse_compositional_algorithm_b200.build --force``)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inverse_compositional_algorithm_b200 import _native, synthetic
from inverse_compositional_algorithm_b200.transformation import TransformType

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nscales = int(sys.argv[2]) if len(sys.argv) > 2 else 5
H = W = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
CH = int(sys.argv[4]) if len(sys.argv) > 4 else 3
t = TransformType.HOMOGRAPHY
pairs = [synthetic.make_pair(i, H, W, CH, t) for i in range(min(B, 4))]
I1 = np.stack([pairs[i % len(pairs)][0] for i in range(B)]); I2 = np.stack([pairs[i % len(pairs)][1] for i in range(B)])
plan = _native.Plan(batch=B, height=H, width=W, channels=CH, nscales=nscales, nu=0.5, transform_type=t.value,
                    robust_type=3, robust_loop=True, lambda_=0.0, tol=1e-3, max_iter=30, delta=10, nanifoutside=True)
plan.run_host(I1, I2)
plan.debug_timeline(True)
plan.run_host(I1, I2)
tl = plan.debug_timeline(True, fetch=True)[:-1].astype(np.float64)
s = tl.sum(axis=0)
names = ["tiles", "tiles without box", "warp steps", "slow pixels", "clk wait mbar", "clk consumer total", "clk try_issue",
         "clk flush", "clk wait desc", "clk chunk epilogue", "chunks"]
for i, n in enumerate(names):
    print(f"{n:22s} {s[i]:.4g}")
tot = s[5]
print("fractions of consumer time: mbar %.3f issue %.3f flush %.3f desc %.3f epilogue %.3f" % (s[4] / tot, s[6] / tot, s[7] / tot, s[8] / tot, s[9] / tot))
print("clk per warp step %.1f ; slow pixels per warp step %.3f" % (tot / max(s[2], 1), s[3] / max(s[2], 1)))
