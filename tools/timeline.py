"""Per-CTA timeline of one iterate launch (profiling hook; needs a library built with ICA_TIMELINE=1:
``ICA_TIMELINE=1 python -m inverse_compositional_algorithm_b200.build --force``).  Runs a 1-pair registration limited
to a given number of launches so that the last launch is the one of interest."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inverse_compositional_algorithm_b200 import _native, synthetic
from inverse_compositional_algorithm_b200.transformation import TransformType

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nscales = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H = W = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 3
CH = int(sys.argv[5]) if len(sys.argv) > 5 else 3
RT = int(sys.argv[6]) if len(sys.argv) > 6 else 3
TT = sys.argv[7] if len(sys.argv) > 7 else 'HOMOGRAPHY'
t = TransformType[TT]
pairs = [synthetic.make_pair(i, H, W, CH, t, max_shift=1.0) for i in range(min(B, 2))]
I1 = np.stack([pairs[i % len(pairs)][0] for i in range(B)]); I2 = np.stack([pairs[i % len(pairs)][1] for i in range(B)])
plan = _native.Plan(batch=B, height=H, width=W, channels=CH, nscales=nscales, nu=0.5, transform_type=t.value,
                    robust_type=RT, robust_loop=RT != 0, lambda_=0.0, tol=1e-9, max_iter=max_iter, delta=10, nanifoutside=True)
plan.debug_timeline(True)
plan.run_host(I1, I2)
plan.run_host(I1, I2)
tl = plan.debug_timeline(True, fetch=True)
solve_row = tl[-1]; tl = tl[:-1]
act = tl[tl[:, 0] > 0]
t0 = act[:, 0].min()
names = ["start", "first_tile_ready", "tiles_done"]
print("CTAs with work:", len(act))
for i, nm in enumerate(names):
    col = act[:, i]
    col = col[col > 0]
    if len(col):
        print(f"{nm:18s} n={len(col):4d}  min={(col.min()-t0)/1e3:8.2f} us  median={(np.median(col)-t0)/1e3:8.2f}  max={(col.max()-t0)/1e3:8.2f}")

end = act[:, 13]; end = end[end > 0] - t0
print("block_end percentiles (us):", [round(float(np.percentile(end, q)) / 1e3, 1) for q in (0, 10, 25, 50, 75, 90, 99, 100)])
print("items per block:", np.unique(act[:, 14], return_counts=True))

sr = solve_row
if sr[0] > 0:   # stamps of the stand-alone solve kernel (ICA_NO_FUSE=1); the fused solve runs inside the iterate kernel
    print("solve kernel block 0 (us since its start): staged %.2f sums %.2f assembled %.2f gj %.2f updated %.2f" % tuple((sr[i] - sr[0]) / 1e3 for i in range(1, 6)))
    print("iterate max end -> solve start: %.2f us" % ((sr[0] - act[:, 13].max()) / 1e3))

# cycle accumulators of the LAST launch (consumer warp 0 / producer lane 0), as fractions of the CTA's consumer lifetime
tot = act[:, 6].astype(np.float64)
ok = tot > 0
if ok.any():
    f = lambda c: float(np.mean(act[ok, c] / tot[ok]))
    print("consumer warp 0: waiting for a full stage %.1f%%, chunk epilogues %.1f%%, border fills %.1f%% of its lifetime; tiles per CTA median %d" %
          (100 * f(3), 100 * f(4), 100 * f(5), int(np.median(act[ok, 15]))))
    print("producer: fetching/decoding work items %.1f%%, waiting for an empty stage %.1f%%, window projection %.1f%%, tile control %.1f%%, "
          "row plans + copy issue %.1f%% of the consumer lifetime" % (100 * f(7), 100 * f(8), 100 * f(10), 100 * f(11), 100 * f(12)))
