"""Load the UNMODIFIED reference (/root/reference/src) behind the skimage/imageio
shim.  Only works in the build container (the reference does not travel to the
GPU box); used by ``oracle/make_golden.py`` and by CPU tests that are skipped
when the reference is absent.  Test infrastructure only."""
import importlib
import os
import sys

REFERENCE_SRC = os.environ.get("ICA_REFERENCE_SRC", "/root/reference/src")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "inverse_compositional_algorithm.py"))


def load():
    """Returns a dict of the reference's modules keyed by their short names."""
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    for path in (_SHIM, REFERENCE_SRC):
        if path not in sys.path:
            sys.path.insert(0, path)
    names = {
        "ica": "inverse_compositional_algorithm",
        "tr": "transformation",
        "de": "derivatives",
        "io": "image_optimisation",
        "bi": "bicubic_interpolation",
        "zm": "zoom",
        "cts": "constants",
        "cfh": "configuration_handler",
    }
    return {short: importlib.import_module(full) for short, full in names.items()}
