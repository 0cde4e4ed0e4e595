"""Import shim so the UNMODIFIED reference sources (/root/reference/src) can be
imported in a container without scikit-image.  Test infrastructure only; the
arithmetic lives in ``oracle/skimage_restated.py``."""
import numpy as _np

from . import transform, filters, exposure  # noqa: F401


def img_as_ubyte(image):  # imported by the reference, never called on the path
    return _np.clip(_np.round(_np.asarray(image) * 255.0), 0, 255).astype(_np.uint8)
