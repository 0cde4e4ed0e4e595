import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from inverse_compositional_algorithm_b200 import synthetic
from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
from inverse_compositional_algorithm_b200.transformation import TransformType
types = [TransformType.SIMILARITY, TransformType.AFFINITY] * 3
pairs = [synthetic.make_pair(40 + i, 120, 160, 1, t, max_shift=4.0, margin=32) for i, t in enumerate(types)]
I1 = np.stack([a for a, _, _ in pairs]); I2 = np.stack([b for _, b, _ in pairs])
p, err, iters = register_batch(I1, I2, types, nscales=3, delta=5)
for rep in range(2):
    for i, t in enumerate(types):
        ps, es, its = register_batch(I1[i:i + 1], I2[i:i + 1], t, nscales=3, delta=5)
        print(rep, i, iters[i], its[0], np.abs(p[i] - ps[0]).max(), err[i], es[0])
