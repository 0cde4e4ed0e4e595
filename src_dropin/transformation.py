"""Top-level ``transformation`` for the reference's binding idiom: its callers do
``sys.path.append("../src/")`` and then ``import transformation`` / ``from transformation import ...``
(``/root/reference/test/inverse_compositional_algorithm_robust.ipynb:49-51``, ``test/test_derivatives.py:7-9``).
Pointing that path at this directory instead binds the same names to the B200 package (mirror of ``src/transformation.py``)."""
from _b200_path import PACKAGE as _PACKAGE  # noqa: F401  (puts the repository root on sys.path)
from inverse_compositional_algorithm_b200.transformation import *  # noqa: F401,F403,E402
from inverse_compositional_algorithm_b200.transformation import (TransformType, update_transform, project, params2matrix, matrix2params, transform_image)  # noqa: F401,E402
