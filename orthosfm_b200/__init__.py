"""orthosfm_b200 -- B200-native exhaustive pairwise feature matcher for OrthoSfM.

Only the one hot path is here: what runs behind sfm::MatchingBase in the reference
(src/mve/sfm/exhaustive_matching.{h,cc}).  The CUDA kernels and the C ABI live in
``csrc/`` (built into ``csrc/libosfm_match.so``); this package is the Python mirror of
the reference interface used by the tests and the benchmark.
"""
from ._lib import KIND_SIFT_U8, KIND_SURF_S8, MatcherError  # noqa: F401
from .matcher import (ExhaustiveMatching, PackedViews, FeatureSet, Matching, MatchingBase,  # noqa: F401
                      TwoViewOptions, Viewport, TWO_VIEW_OK, TWO_VIEW_SKIPPED,
                      TWO_VIEW_LOWRES_REJECTED, TWO_VIEW_TOO_FEW_MATCHES, TWO_VIEW_TOO_FEW_INLIERS,
                      ransac_draw_samples)

__all__ = ["ExhaustiveMatching", "PackedViews", "FeatureSet", "Matching", "MatchingBase", "Viewport",
           "MatcherError", "KIND_SIFT_U8", "KIND_SURF_S8", "TwoViewOptions", "TWO_VIEW_OK",
           "TWO_VIEW_SKIPPED", "TWO_VIEW_LOWRES_REJECTED", "TWO_VIEW_TOO_FEW_MATCHES", "TWO_VIEW_TOO_FEW_INLIERS",
           "ransac_draw_samples"]
