"""feature_tracker_b200 -- B200-native (sm_100a) sparse KLT tracking and descriptor matching.

The product is libftk_b200.so (hand-written CUDA kernels behind the C ABI of include/ftk_c.h); this package is the
thin host-side mirror of the reference's C++ interface used by the tests and the benchmark.
"""
from . import _capi  # noqa: F401
from .api import (  # noqa: F401
    BriefDescriptor, BriefMatcher, Context, CosineMatcher, DenseOpticalFlow, DescriptorMatcher, DirectMethod, DirectMethodMethod, DirectMethodOptions, DiskMatcher, FeaturePointHarrisDetector, FeaturePointShiTomasDetector, FtkError, ImagePyramidBatch, MatcherOptions, NNFeatureMatcher, OpticalFlow,
    OpticalFlowAffineKlt, OpticalFlowBasicKlt, OpticalFlowLssdKlt, OpticalFlowMethod, OpticalFlowOptions, SuperpointMatcher, TrackStatus,
    brief_pattern, default_context, pack_brief,
)
