"""Iteration statistics of the bench workload for several seeds (which rank's batch is the hard one)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from inverse_compositional_algorithm_b200 import _native, synthetic
from inverse_compositional_algorithm_b200.transformation import TransformType, end_point_error
t = TransformType.HOMOGRAPHY
B = 32
plan = _native.Plan(batch=B, height=1024, width=1024, channels=3, nscales=5, nu=0.5, transform_type=t.value, robust_type=3,
                    robust_loop=True, lambda_=0.0, tol=1e-3, max_iter=30, delta=10, nanifoutside=True, gray_as_rgb=False)
for seed in [int(a) for a in sys.argv[1:]] or [1, 1001]:
    I1, I2, p_gt = synthetic.make_batch_torch(B, 1024, 1024, 3, [t] * B, seed=seed, device="cuda")
    I1 = I1.round_().clamp_(0, 255); I2 = I2.round_().clamp_(0, 255)
    p = torch.zeros((B, 8), dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        p.zero_(); e0.record()
        plan.run_device(I1.data_ptr(), I2.data_ptr(), p.data_ptr(), torch.cuda.current_stream().cuda_stream)
        e1.record(); torch.cuda.synchronize()
    pr, err, iters = plan.results()
    epe = np.array([end_point_error(pr[i], p_gt[i], t, 1024, 1024)[1] for i in range(B)])
    print("seed", seed, "ms", round(e0.elapsed_time(e1), 3), "iters mean (fine->coarse)", iters.mean(0).round(2), "max", iters.max(0),
          "epe max", epe.max().round(4), "n(epe>0.1)", int((epe > 0.1).sum()))
    bad = np.where(iters.max(1) >= 30)[0]
    for i in bad[:6]:
        print("   pair", i, "iters", iters[i], "epe", epe[i].round(4), "p_gt", p_gt[i].round(5))
