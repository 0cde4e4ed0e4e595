"""Top-level ``bicubic_interpolation`` for the reference's binding idiom: its callers do
``sys.path.append("../src/")`` and then ``import bicubic_interpolation`` / ``from bicubic_interpolation import ...``
(``/root/reference/test/inverse_compositional_algorithm_robust.ipynb:49-51``, ``test/test_derivatives.py:7-9``).
Pointing that path at this directory instead binds the same names to the B200 package (mirror of ``src/bicubic_interpolation.py``)."""
from _b200_path import PACKAGE as _PACKAGE  # noqa: F401  (puts the repository root on sys.path)
from inverse_compositional_algorithm_b200.bicubic_interpolation import *  # noqa: F401,F403,E402
from inverse_compositional_algorithm_b200.bicubic_interpolation import (neumann_bc, cubic_interpolation, bicubic_interpolation_array, bicubic_interpolation_point, bicubic_interpolation_image, bicubic_interpolation_skimage)  # noqa: F401,E402
