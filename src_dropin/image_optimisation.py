"""Top-level ``image_optimisation`` for the reference's binding idiom: its callers do
``sys.path.append("../src/")`` and then ``import image_optimisation`` / ``from image_optimisation import ...``
(``/root/reference/test/inverse_compositional_algorithm_robust.ipynb:49-51``, ``test/test_derivatives.py:7-9``).
Pointing that path at this directory instead binds the same names to the B200 package (mirror of ``src/image_optimisation.py``)."""
from _b200_path import PACKAGE as _PACKAGE  # noqa: F401  (puts the repository root on sys.path)
from inverse_compositional_algorithm_b200.image_optimisation import *  # noqa: F401,F403,E402
from inverse_compositional_algorithm_b200.image_optimisation import (RobustErrorFunctionType, rhop, robust_error_function, independent_vector, independent_vector_robust, parametric_solve, steepest_descent_images)  # noqa: F401,E402
