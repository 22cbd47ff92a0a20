"""r3det_b200 — B200 (sm_100a) implementation of R3Det's rotated-geometry hot path behind the
reference's own Python operator API (r3det/ops, r3det/core).  Hand-written CUDA kernels in csrc/,
reached through the C ABI of include/r3geo.h; no CPU fallback.

Module map (reference module -> here):
  r3det/ops/rbbox_geo            -> rbbox_geo         rbbox_iou
  r3det/ops/box_iou_rotated      -> box_iou_rotated   obb_overlaps
  r3det/ops/rnms                 -> rnms              rnms, batched_rnms
  r3det/ops/nms_rotated          -> nms_rotated       obb_nms, obb_batched_nms, poly_nms
  r3det/ops/ml_nms_rotated       -> ml_nms_rotated    ml_nms_rotated
  r3det/core/bbox/iou_calculators-> iou_calculators   RBboxOverlaps2D_v1/v2/v3, rbbox_overlaps_v1/v2/v3
  r3det/core/post_processing     -> bbox_nms_rotated  multiclass_nms_rotated
  r3det/ops/fr                   -> fr                FeatureRefineFunction, feature_refine, FR, FeatureRefineModule
  (mmdet MaxIoUAssigner + calculator, fused; §8f) -> assign   max_iou_assign, max_iou_assign_batched, FusedMaxIoUAssigner
  r3det/core/bbox/coder          -> coder             DeltaXYWHAOBBoxCoder, bbox2delta_v1/2/3, delta2bbox_v1/2/3
  r3det/models/dense_heads (get_bboxes tail, filter_bboxes, refine_bboxes; §8f) -> dense_tail
  r3det/core/bbox/rtransforms    -> rtransforms       poly2obb, obb2poly, obb2hbb, hbb2obb, obb2xyxy, norm_angle, *_np, ...
  r3det/datasets/dota1.py (merge_det, _merge_func, _results2submission; §8f) -> dota_submission
"""
from . import _lib  # noqa: F401
from ._nms_core import pack_keep_records  # noqa: F401
from .assign import FusedMaxIoUAssigner, max_iou_assign, max_iou_assign_batched  # noqa: F401
from .bbox_nms_rotated import multiclass_nms_rotated, multiclass_nms_rotated_batch, multiclass_nms_rotated_padded  # noqa: F401
from .box_iou_rotated import obb_overlaps  # noqa: F401
from .coder import (DeltaXYWHAOBBoxCoder, bbox2delta_v1, bbox2delta_v2, bbox2delta_v3, delta2bbox_v1,  # noqa: F401
                    delta2bbox_v2, delta2bbox_v3)
from .dense_tail import filter_bboxes, get_bboxes, get_bboxes_padded, refine_bboxes, select_decode  # noqa: F401
from .fr import (FR, FeatureRefineFunction, FeatureRefineModule, FeatureRefineMultiFunction, feature_refine,  # noqa: F401
                 feature_refine_multi)
from .iou_calculators import (IOU_CALCULATORS, RBboxOverlaps2D_v1, RBboxOverlaps2D_v2,  # noqa: F401
                              RBboxOverlaps2D_v3, rbbox_overlaps_v1, rbbox_overlaps_v2, rbbox_overlaps_v3)
from .ml_nms_rotated import ml_nms_rotated  # noqa: F401
from .nms_rotated import obb_batched_nms, obb_nms, poly_nms  # noqa: F401
from .rbbox_geo import aligned_iou, pairwise_iou, rbbox_iou  # noqa: F401
from .rnms import batched_rnms, rnms  # noqa: F401
from .rtransforms import (hbb2obb, norm_angle, obb2hbb, obb2poly, obb2poly_np, obb2xyxy, poly2obb, poly2obb_np,  # noqa: F401
                          rbbox2result, rbbox2roi)
from . import dota_submission  # noqa: F401
