"""pcf_b200 -- B200-native (sm_100a) hot path of PointConvFormer behind the reference's own interfaces.

Sub-modules mirror the reference's module names for this path:
    pcf_cuda                  <- cpp_wrappers/cpp_pcf_kernel (python module `pcf_cuda`)
    layer_utils, layers       <- layer_utils.py, layers.py
    model_architecture        <- model_architecture.py
    knn_post_dataloader_utils <- knn_post_dataloader_utils.py
    common_util               <- util/common_util.py (compute_knn_inverse, replace_batchnorm)
    grid_subsampling          <- cpp_wrappers/cpp_subsampling + datasetCommon.subsample
All compute goes through libpcf_b200.so (include/pcf_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["pcf_cuda", "layer_utils", "layers", "model_architecture", "knn_post_dataloader_utils",
           "common_util", "grid_subsampling", "synthetic"]
