"""B200-native SIFT detection engine: host mirror of the reference's detection path.

`sift` mirrors src/sift.js, `background` mirrors background.js (+ fused detect);
both forward to libsift_b200.so (CUDA, sm_100a) through the C ABI declared in
include/sift_b200.h.  There is no CPU path in this package.
"""
from . import _lib, fixtures  # noqa: F401
from ._lib import (CANDIDATE_DTYPE, KEYPOINT_DTYPE, Params, SiftError, default_params)  # noqa: F401
from .engine import Engine, default_engine  # noqa: F401
from . import sift, background, mosaic  # noqa: F401
from .sift import (SIFT_blurMatrix2DChunk, SIFT_findExtremas, SIFT_generateGradientVector,  # noqa: F401
                   SIFT_generateHessianMatrix, SIFT_subtractMatrix2DChunk)
from .background import (computeDifferenceOfGaussians, computeGaussianScaleSpace, detect,  # noqa: F401
                         findCandidateKeypoints, refineCandidateKeypoints)

__version__ = "0.1.0"
