"""lane_tracker_b200 -- B200-native implementation of the lane_tracker per-frame hot path.

The tracker classes need the in-tree CUDA library (liblane_tracker_b200.so) and a CUDA device;
importing them without either raises.  ``lane_tracker_b200.synth`` (test/bench data) and
``lane_tracker_b200.utils`` (calibration loaders) are plain NumPy.
"""
from .utils import create_split_view, load_camera_calib, load_warp_params  # noqa: F401

__all__ = ["LaneTracker", "BatchedLaneTracker", "HostPipeline", "DevicePipeline", "GraphedProcess", "bilateral_adaptive_threshold",
           "load_camera_calib", "load_warp_params", "create_split_view"]


def __getattr__(name):
    if name in ("LaneTracker", "BatchedLaneTracker", "HostPipeline", "DevicePipeline", "GraphedProcess", "make_params", "RESULT_DTYPE",
                "bilateral_adaptive_threshold"):
        from . import tracker
        return getattr(tracker, name)
    raise AttributeError(name)
