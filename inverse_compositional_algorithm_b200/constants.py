"""Algorithm constants; same names and values as the reference's ``src/constants.py:1-6``.
The native library carries its own copies (``csrc/ica_common.cuh``); a test checks they agree."""
MAX_ITER = 30
LAMBDA_0 = 80.0
LAMBDA_N = 5.0
LAMBDA_RATIO = 0.9

ZOOM_SIGMA_ZERO = 0.6
