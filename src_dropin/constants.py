"""Top-level ``constants`` for the reference's binding idiom: its callers do
``sys.path.append("../src/")`` and then ``import constants`` / ``from constants import ...``
(``/root/reference/test/inverse_compositional_algorithm_robust.ipynb:49-51``, ``test/test_derivatives.py:7-9``).
Pointing that path at this directory instead binds the same names to the B200 package (mirror of ``src/constants.py``)."""
from _b200_path import PACKAGE as _PACKAGE  # noqa: F401  (puts the repository root on sys.path)
from inverse_compositional_algorithm_b200.constants import *  # noqa: F401,F403,E402
from inverse_compositional_algorithm_b200.constants import (MAX_ITER, LAMBDA_0, LAMBDA_N, LAMBDA_RATIO, ZOOM_SIGMA_ZERO)  # noqa: F401,E402
