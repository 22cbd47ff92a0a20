"""multiclass_nms_rotated — mirror of r3det/core/post_processing/bbox_nms_rotated.py:7-131.

Same signature, same candidate order (row-major over (box, class) of scores > score_thr, :98-103), same
per-variant output order and truncation:
  v1   batched_rnms      keep in ascending candidate index, dets[:max_num] in that order (:111,127-129)
  v3   obb_batched_nms   keep in descending score, dets[:max_num] = top scores (:113)
  v2   ml_nms_rotated    keep in score order; `keep.size(0) > max_num` re-sort and slice (:115-125), including
                         the reference's max_num=-1 quirk that drops the lowest-scoring detection
  mmcv mmcv.ops.nms_rotated(bboxes, scores, iou_thr, labels) — third-party; served by the v2 geometry with
                         labels (same lineage, +a rotation), dets/keep in score order, `return_inds` honoured.
`nms` is an mmcv ConfigDict or dict with `iou_thr` and optional `type` (default 'v1', :43)."""
import torch

from .ml_nms_rotated import ml_nms_rotated
from .nms_rotated import obb_batched_nms
from .rnms import batched_rnms


def _get(nms, key, default=None):
    if isinstance(nms, dict):
        return nms.get(key, default)
    return getattr(nms, key, default)


def multiclass_nms_rotated(multi_bboxes, multi_scores, score_thr, nms, max_num=-1, score_factors=None,
                           return_inds=False):
    num_classes = multi_scores.size(1) - 1
    if multi_bboxes.shape[1] > 5:
        bboxes = multi_bboxes.view(multi_scores.size(0), -1, 5)
    else:
        bboxes = multi_bboxes[:, None].expand(multi_scores.size(0), num_classes, 5)
    scores = multi_scores[:, :-1]
    nms_version = _get(nms, 'type', 'v1')
    iou_thr = _get(nms, 'iou_thr')

    if nms_version == 'mmcv':
        labels = torch.arange(num_classes, dtype=torch.long, device=scores.device)
        labels = labels.view(1, -1).expand_as(scores)
        bboxes = bboxes.reshape(-1, 5)
        scores = scores.reshape(-1)
        labels = labels.reshape(-1)
        valid_mask = scores > score_thr
        if score_factors is not None:
            score_factors = score_factors.view(-1, 1).expand(multi_scores.size(0), num_classes)
            scores = scores * score_factors.reshape(-1)
        inds = valid_mask.nonzero(as_tuple=False).squeeze(1)
        bboxes, scores, labels = bboxes[inds], scores[inds], labels[inds]
        if bboxes.numel() == 0:
            dets = torch.cat([bboxes, scores[:, None]], -1)
            return (dets, labels, inds) if return_inds else (dets, labels)
        keep = ml_nms_rotated(bboxes, scores, labels, iou_thr)
        dets = torch.cat([bboxes[keep], scores[keep][:, None]], -1)
        if max_num > 0:
            dets = dets[:max_num]
            keep = keep[:max_num]
        return (dets, labels[keep], keep) if return_inds else (dets, labels[keep])

    valid_mask = scores > score_thr
    bboxes = bboxes[valid_mask]
    if score_factors is not None:
        scores = scores * score_factors[:, None]
    scores = scores[valid_mask]
    labels = valid_mask.nonzero(as_tuple=False)[:, 1]

    if bboxes.numel() == 0:
        bboxes = multi_bboxes.new_zeros((0, 6))
        labels = multi_bboxes.new_zeros((0, ), dtype=torch.long)
        return bboxes, labels

    if nms_version == 'v1':
        dets, keep = batched_rnms(bboxes, scores, labels, iou_thr)
    elif nms_version == 'v3':
        dets, keep = obb_batched_nms(bboxes, scores, labels, iou_thr)
    elif nms_version == 'v2':
        keep = ml_nms_rotated(bboxes, scores, labels, iou_thr)
        bboxes = bboxes[keep]
        scores = scores[keep]
        labels = labels[keep]
        if keep.size(0) > max_num:
            _, inds = scores.sort(descending=True)
            inds = inds[:max_num]
            bboxes = bboxes[inds]
            scores = scores[inds]
            labels = labels[inds]
        return torch.cat([bboxes, scores[:, None]], 1), labels
    else:
        raise KeyError(f'unknown rotated nms type {nms_version!r}')

    if max_num > 0:
        dets = dets[:max_num]
        keep = keep[:max_num]
    return dets, labels[keep]
