"""Locates the B200 package for the flat drop-in modules of this directory."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PACKAGE = "inverse_compositional_algorithm_b200"
