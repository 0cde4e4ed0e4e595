"""Where does the end-to-end time go?  H2D bandwidth from pinned memory, one plan's run_host time, N plans on N threads."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from inverse_compositional_algorithm_b200 import _native, synthetic
from inverse_compositional_algorithm_b200.transformation import TransformType

t = TransformType.HOMOGRAPHY
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 32
NP = int(sys.argv[2]) if len(sys.argv) > 2 else 4
I1, I2, _ = synthetic.make_batch_device(NB, 1024, 1024, 3, [t] * NB, seed=1, device="cuda")
h1 = I1.round().clamp(0, 255).to(torch.uint8).cpu().pin_memory()
h2 = I2.round().clamp(0, 255).to(torch.uint8).cpu().pin_memory()
d = torch.empty_like(h1, device="cuda")
for _ in range(3):
    d.copy_(h1, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(h1, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D pinned {h1.numel()/1e6:.0f} MB: {dt*1e3:.2f} ms = {h1.numel()/dt/1e9:.1f} GB/s")

plans = []
for i in range(NP):
    pl = _native.Plan(batch=NB, height=1024, width=1024, channels=3, nscales=5, nu=0.5, transform_type=t.value, robust_type=3,
                      robust_loop=True, lambda_=0.0, tol=1e-3, max_iter=30, delta=10, nanifoutside=True, gray_as_rgb=False)
    plans.append(dict(plan=pl, p=np.zeros((NB, 8)), err=np.zeros(NB), it=np.zeros((NB, 5), dtype=np.int32)))

def work(hv, n):
    for _ in range(n):
        hv["p"][:] = 0
        hv["plan"].run_host_ptrs(h1.data_ptr(), h2.data_ptr(), _native.DTYPE_U8, hv["p"], hv["err"], hv["it"])

work(plans[0], 2)
t0 = time.perf_counter(); work(plans[0], 5); dt = (time.perf_counter() - t0) / 5
print(f"one plan, one thread: {dt*1e3:.2f} ms per {NB} pairs = {NB/dt:.0f} pairs/s; device span of the last call {plans[0]['plan'].last_host_run_ms():.2f} ms")
for n in (2, 3, 4, 6, 8)[: max(1, NP - 1)]:
    if n > NP: break
    ths = [threading.Thread(target=work, args=(plans[i], 2)) for i in range(n)]
    [x.start() for x in ths]; [x.join() for x in ths]
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(plans[i], 5)) for i in range(n)]
    [x.start() for x in ths]; [x.join() for x in ths]
    dt = time.perf_counter() - t0
    print(f"{n} plans on {n} threads: {n*5*NB/dt:.0f} pairs/s ({dt/5*1e3:.2f} ms per round of {n*NB} pairs)")
