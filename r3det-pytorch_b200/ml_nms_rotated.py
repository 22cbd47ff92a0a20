"""v2 multi-label rotated NMS — mirror of ml_nms_rotated_cuda.ml_nms_rotated
(r3det/ops/ml_nms_rotated/src/nms_rotated.h:23-38): boxes of different labels never suppress each other
(box_iou_rotated_utils.h:317-322), rotation +a, keep list in descending-score order."""
from ._nms_core import nms_device, to_cuda_input


def ml_nms_rotated(dets, scores, labels, iou_threshold):
    """dets (K,5), scores (K,), labels (K,) -> kept indices (int64), score order."""
    d, _, was_host = to_cuda_input(dets, None, "dets")
    keep, num = nms_device(d, scores.to(d.device), iou_threshold, "v2", labels=labels.to(d.device), inclusive=was_host)
    keep = keep[:int(num.item())]
    return keep.cpu() if was_host else keep
