"""yolov10-3d_b200: B200-native (sm_100a) head decode, NMS-free top-k and dual task-aligned assignment for
YOLOv10 / YOLOv10-3D, behind the reference's own Python call signatures.

The directory name carries a hyphen, so import it through the root-level alias module ``yolov10_3d_b200``.
Everything computes in liby3d_b200.so (hand-written CUDA, C ABI in include/y3d.h); there is no CPU or PyTorch
fallback -- a missing library raises at the first call.
"""
from . import _lib  # noqa: F401
from . import dist, head, kitti, loss, loss3d, ops, tal  # noqa: F401
from ._lib import Y3DError, lib  # noqa: F401
from .head import (V10DetectDecoder, detect3d_decode, detect3d_postprocess, detect_inference, extract_patches,  # noqa: F401
                   inference_forward_feat, scatter_candidates, select_candidates, v10detect_export_forward)
from .loss import v8DetectionLoss, v10DetectLoss  # noqa: F401
from .loss3d import DDDetectionLoss, DetectLoss3d  # noqa: F401
from .ops import v10_3Dpostprocess, v10postprocess, xywh2xyxy  # noqa: F401
from .tal import TaskAlignedAssigner, TaskAlignedAssigner3d, make_anchors  # noqa: F401

__version__ = "0.1.0"
