"""Top-level ``configuration_handler`` for the reference's binding idiom: its callers do
``sys.path.append("../src/")`` and then ``import configuration_handler`` / ``from configuration_handler import ...``
(``/root/reference/test/inverse_compositional_algorithm_robust.ipynb:49-51``, ``test/test_derivatives.py:7-9``).
Pointing that path at this directory instead binds the same names to the B200 package (mirror of ``src/configuration_handler.py``)."""
from _b200_path import PACKAGE as _PACKAGE  # noqa: F401  (puts the repository root on sys.path)
from inverse_compositional_algorithm_b200.configuration_handler import *  # noqa: F401,F403,E402
from inverse_compositional_algorithm_b200.configuration_handler import (create_config_file, read_config_file)  # noqa: F401,E402
