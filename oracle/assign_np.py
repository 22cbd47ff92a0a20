"""TEST INFRASTRUCTURE ONLY — numpy restatement of mmdet-2.19 `MaxIoUAssigner` (the consumer of the IoU calculators on the
training path: r3det/models/dense_heads/rotate_anchor_head.py:220-228 builds it from `train_cfg.assigner` and calls
`assign(anchors, gt_bboxes, gt_bboxes_ignore, gt_labels)`).

PROVENANCE — PARITY UNPINNED UPSTREAM.  mmdet is a third-party dependency of the reference (`mmdet==2.19.0`, README.md:13); its
source is not under /root/reference and cannot be installed here (no network).  The functions below restate
`mmdet/core/bbox/assigners/max_iou_assigner.py` (v2.19.0) — `assign` (:81-144) and `assign_wrt_overlaps` (:146-212) — from the
published algorithm as recalled in SURVEY.md A6; line numbers are approximate.  What pins them: the known-answer example in
that file's own docstring (two boxes, one GT -> gt_inds [1, 0]) and hand-worked cases in tests/test_oracle_assign.py
(thresholds, low-quality matching with and without gt_max_assign_all, ties, ignore regions, tuple neg_iou_thr, empty sides).
The product never imports this module."""
import numpy as np


def assign_wrt_overlaps(overlaps, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, match_low_quality=True, gt_max_assign_all=True):
    """overlaps (G, A) -> (assigned_gt_inds (A,) int64: -1 ignore / 0 background / i + 1 matched to GT i, max_overlaps (A,))."""
    overlaps = np.asarray(overlaps, np.float32)
    G, A = overlaps.shape
    assigned = np.full((A,), -1, np.int64)
    if G == 0 or A == 0:
        if G == 0:
            assigned[:] = 0
        return assigned, np.zeros((A,), np.float32)
    max_ov, argmax = overlaps.max(0), overlaps.argmax(0)                # per anchor (first maximum, like torch.max)
    gt_max, gt_argmax = overlaps.max(1), overlaps.argmax(1)             # per GT
    if isinstance(neg_iou_thr, (tuple, list)):
        assert len(neg_iou_thr) == 2
        assigned[(max_ov >= neg_iou_thr[0]) & (max_ov < neg_iou_thr[1])] = 0
    else:
        assigned[(max_ov >= 0) & (max_ov < neg_iou_thr)] = 0
    pos = max_ov >= pos_iou_thr
    assigned[pos] = argmax[pos] + 1
    if match_low_quality:
        # later GTs overwrite earlier ones; an anchor already positive for another GT is re-assigned (known mmdet behaviour)
        for i in range(G):
            if gt_max[i] >= min_pos_iou:
                if gt_max_assign_all:
                    assigned[overlaps[i, :] == gt_max[i]] = i + 1
                else:
                    assigned[gt_argmax[i]] = i + 1
    return assigned, max_ov


def assign(overlaps, ignore_iof=None, ignore_iof_thr=-1, **kw):
    """`assign` around assign_wrt_overlaps: columns whose best IoF with an ignore region exceeds ignore_iof_thr are set to -1
    first (ignore_iof: (A,) = max over ignore boxes of iof(anchor, ignore) for ignore_wrt_candidates=True, or of
    iof(ignore, anchor) otherwise — the caller computes it with the calculator under test)."""
    overlaps = np.array(overlaps, np.float32, copy=True)
    if ignore_iof_thr > 0 and ignore_iof is not None and overlaps.shape[1] > 0 and np.size(ignore_iof) > 0:
        overlaps[:, np.asarray(ignore_iof) > ignore_iof_thr] = -1
    return assign_wrt_overlaps(overlaps, **kw)
