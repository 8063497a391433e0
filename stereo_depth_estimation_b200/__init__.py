"""B200-native (sm_100a) train / infer step of the stereo U-Net disparity model of
sdfgeoff/stereo_depth_estimation.  Host side in Python over a C-ABI CUDA library
(include/sdn.h -> libsdn_b200.so); see DESIGN.md and INTEGRATION.md."""
from .model import ConvBlock, StereoUNet, load_state_dict_compat  # noqa: F401

from .pipeline import SourcePrefetcher  # noqa: F401
from .live import LivePipeline  # noqa: F401

__all__ = ["StereoUNet", "ConvBlock", "load_state_dict_compat", "SourcePrefetcher", "LivePipeline"]
