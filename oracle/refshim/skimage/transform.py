"""`skimage.transform` surface used by the reference: warp, rescale and the four
matrix transform classes (only `.params` and `.inverse` are touched:
bi.py:175-199, tr.py:292-316)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import skimage_restated as _sr  # noqa: E402


class ProjectiveTransform:
    def __init__(self, matrix=None):
        self.params = np.eye(3) if matrix is None else np.array(matrix, dtype=np.float64)

    @property
    def inverse(self):
        return type(self)(matrix=np.linalg.inv(self.params))


class AffineTransform(ProjectiveTransform):
    pass


class SimilarityTransform(ProjectiveTransform):
    pass


class EuclideanTransform(ProjectiveTransform):
    def __init__(self, matrix=None, rotation=None, translation=None):
        if matrix is not None:
            super().__init__(matrix=matrix)
            return
        rot = 0.0 if rotation is None else rotation
        tx, ty = (0.0, 0.0) if translation is None else translation
        super().__init__(matrix=[[np.cos(rot), -np.sin(rot), tx],
                                 [np.sin(rot), np.cos(rot), ty],
                                 [0.0, 0.0, 1.0]])


def warp(image, inverse_map, map_args=None, output_shape=None, order=None, mode="constant",
         cval=0.0, clip=True, preserve_range=False):
    assert preserve_range and not map_args and output_shape is None
    if order is None:
        order = 1  # skimage default for non-bool input
    return _sr.warp(image, inverse_map.params, order=order, mode=mode, cval=cval, clip=clip)


def rescale(image, scale, order=None, mode="reflect", cval=0, clip=True, preserve_range=False,
            anti_aliasing=None, anti_aliasing_sigma=None, *, channel_axis=None):
    assert preserve_range and channel_axis == 2 and anti_aliasing_sigma is None
    return _sr.rescale(image, scale, order=order, mode=mode, cval=cval, clip=clip,
                       anti_aliasing=bool(anti_aliasing))
