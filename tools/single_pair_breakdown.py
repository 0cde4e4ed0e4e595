"""Where the time of ONE 1024^2 RGB homography registration goes: device-resident loop (graph), pyramid, K2 sum, gaps."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from inverse_compositional_algorithm_b200 import _native, synthetic
from inverse_compositional_algorithm_b200.transformation import TransformType

t = TransformType.HOMOGRAPHY
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
I1, I2, p_gt = synthetic.make_pair(7, H, W, 3, t)
d1 = torch.from_numpy(np.round(I1)[None]).cuda().contiguous(); d2 = torch.from_numpy(np.round(I2)[None]).cuda().contiguous()
p = torch.zeros(1, 8, dtype=torch.float64, device="cuda")
for timing in (0, 1, 2):
    plan = _native.Plan(batch=1, height=H, width=W, channels=3, nscales=5, nu=0.5, transform_type=t.value, robust_type=3,
                        robust_loop=True, lambda_=0.0, tol=1e-3, max_iter=30, delta=10, nanifoutside=True)
    if timing:
        plan.enable_timing(timing)
    s = torch.cuda.current_stream().cuda_stream
    ts = []
    for rep in range(8):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run_device(d1.data_ptr(), d2.data_ptr(), p.data_ptr(), s)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = plan.results()
    iters = res[2][0]
    line = f"timing mode {timing}: run_device median {np.median(ts[2:]):.3f} ms; iterations per scale (fine->coarse) {iters.tolist()} total {int(iters.sum())}"
    if timing:
        tm = plan.timing()
        line += f"; iterate kernels {tm['iterate_ms']:.3f} ms over {tm['iterate_launches']} launches; pyramid {tm['pyramid_ms']:.3f} ms over {tm['pyramid_launches']} launches"
    print(line)
    plan.close()
