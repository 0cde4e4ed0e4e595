"""Top-level ``inverse_compositional_algorithm`` for the reference's binding idiom: its callers do
``sys.path.append("../src/")`` and then ``import inverse_compositional_algorithm`` / ``from inverse_compositional_algorithm import ...``
(``/root/reference/test/inverse_compositional_algorithm_robust.ipynb:49-51``, ``test/test_derivatives.py:7-9``).
Pointing that path at this directory instead binds the same names to the B200 package (mirror of ``src/inverse_compositional_algorithm.py``)."""
from _b200_path import PACKAGE as _PACKAGE  # noqa: F401  (puts the repository root on sys.path)
from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import *  # noqa: F401,F403,E402
from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (inverse_compositional_algorithm, robust_inverse_compositional_algorithm, pyramidal_inverse_compositional_algorithm, register_batch, register_batch_device, PyramidalInverseCompositional)  # noqa: F401,E402
