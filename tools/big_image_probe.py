import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from inverse_compositional_algorithm_b200 import synthetic
from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch_device
from inverse_compositional_algorithm_b200.transformation import TransformType, end_point_error
t = TransformType.HOMOGRAPHY
for n, c in ((8192, 3), (4096, 3), (6000, 1)):
    I1, I2, p_gt = synthetic.make_batch_device(1, n, n, c, t, seed=3, device="cuda", max_lin=0.002)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p, err, iters = register_batch_device(I1, I2, t, nscales=6, robust_type=3, delta=10)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    epe = end_point_error(p[0].cpu().numpy(), p_gt[0], t, n, n)[1]
    print(n, c, "ms", round(dt * 1e3, 2), "iters", iters[0].tolist(), "EPE vs ground truth", round(float(epe), 4))
    del I1, I2
    torch.cuda.empty_cache()
