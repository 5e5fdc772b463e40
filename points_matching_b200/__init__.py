"""points_matching_b200 -- B200-native (sm_100a) descriptor matching and epipolar geometry.

Host-side mirror of the OpenCV call surface used by the reference program
(/root/reference/Points Matching/main.cpp:43-46, 49-69, 89-91, 95-98, 127-132) over the
C ABI of include/pm.h.  All arithmetic runs in libpm.so's CUDA kernels.
"""
from ._lib import DMATCH, SO_PATH, build  # noqa: F401
from .api import (  # noqa: F401
    FM_7POINT, FM_8POINT, FM_LMEDS, FM_RANSAC, METRIC_SAMPSON, METRIC_SYMEPI, NORM_HAMMING, NORM_L2,
    BFMatcher, Context, PMError, computeCorrespondEpilines, default_context, findFundamentalMat,
    keypoints_convert, minmax_filter, ratio_test)
