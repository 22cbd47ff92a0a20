"""CPU: hand-worked known answers for the numpy restatement of mmdet's MaxIoUAssigner (oracle/assign_np.py).  mmdet itself is
not available (parity unpinned upstream); these cases pin the restatement to the algorithm's documented behaviour."""
import numpy as np

from oracle import assign_np as A


def test_mmdet_docstring_example():
    # max_iou_assigner.py docstring: bboxes [[0,0,10,10],[10,10,20,20]], gt [[0,0,10,9]] with MaxIoUAssigner(0.5, 0.5)
    # -> IoU 0.9 and 0 -> expected_gt_inds [1, 0]
    ov = np.array([[0.9, 0.0]], np.float32)
    got, mx = A.assign_wrt_overlaps(ov, 0.5, 0.5)
    assert got.tolist() == [1, 0] and mx.tolist() == [np.float32(0.9), 0.0]


def test_thresholds_and_gap():
    ov = np.array([[0.70, 0.45, 0.30, 0.0],
                   [0.10, 0.55, 0.35, 0.0]], np.float32)
    got, _ = A.assign_wrt_overlaps(ov, 0.5, 0.4, match_low_quality=False)
    assert got.tolist() == [1, 2, 0, 0]                       # 0.35 < neg -> background
    got, _ = A.assign_wrt_overlaps(ov, 0.6, 0.4, match_low_quality=False)
    assert got.tolist() == [1, -1, 0, 0]                      # 0.55 in the gap [0.4, 0.6) -> ignore (-1)
    got, _ = A.assign_wrt_overlaps(ov, 0.6, (0.2, 0.4), match_low_quality=False)
    assert got.tolist() == [1, -1, 0, -1]                     # tuple neg: [0.2, 0.4) only; 0.0 stays -1


def test_low_quality_matching():
    ov = np.array([[0.30, 0.30, 0.10],
                   [0.20, 0.45, 0.45]], np.float32)
    # no pair reaches pos 0.5; GT0's best 0.30 is shared by anchors 0 and 1, GT1's best 0.45 by anchors 1 and 2
    got, _ = A.assign_wrt_overlaps(ov, 0.5, 0.4, min_pos_iou=0.0, gt_max_assign_all=True)
    assert got.tolist() == [1, 2, 2]                          # anchor 1: GT1 (the later GT) overwrites GT0
    got, _ = A.assign_wrt_overlaps(ov, 0.5, 0.4, min_pos_iou=0.0, gt_max_assign_all=False)
    assert got.tolist() == [1, 2, -1]                         # only the FIRST maximum of each GT; anchor 2 (0.45) sits in the gap [0.4, 0.5)
    got, _ = A.assign_wrt_overlaps(ov, 0.5, 0.46, min_pos_iou=0.0, gt_max_assign_all=False)
    assert got.tolist() == [1, 2, 0]
    got, _ = A.assign_wrt_overlaps(ov, 0.5, 0.4, min_pos_iou=0.4, gt_max_assign_all=True)
    assert got.tolist() == [0, 2, 2]                          # GT0's best 0.30 < min_pos_iou: not matched


def test_zero_row_quirk_and_empty_sides():
    # a GT that overlaps nothing has gt_max = 0; with min_pos_iou = 0 `overlaps[i] == 0` selects EVERY zero column (mmdet quirk)
    ov = np.array([[0.0, 0.0, 0.0],
                   [0.6, 0.0, 0.2]], np.float32)
    got, _ = A.assign_wrt_overlaps(ov, 0.5, 0.4, min_pos_iou=0.0, gt_max_assign_all=True)
    assert got.tolist() == [2, 1, 1]                          # GT0 grabs ALL three anchors (its row is 0 everywhere), then GT1 takes anchor 0 back
    got, _ = A.assign_wrt_overlaps(ov, 0.5, 0.4, min_pos_iou=0.0, gt_max_assign_all=False)
    assert got.tolist() == [2, 0, 0]                          # first maximum of the zero row is anchor 0, re-taken by GT1
    got, mx = A.assign_wrt_overlaps(np.zeros((0, 3), np.float32), 0.5, 0.4)
    assert got.tolist() == [0, 0, 0] and mx.tolist() == [0, 0, 0]
    got, mx = A.assign_wrt_overlaps(np.zeros((2, 0), np.float32), 0.5, 0.4)
    assert got.shape == (0,)


def test_ignore_regions():
    ov = np.array([[0.70, 0.60, 0.10, 0.45]], np.float32)
    iof = np.array([0.0, 0.8, 0.6, 0.2], np.float32)          # anchors 1 and 2 sit inside an ignore region
    got, mx = A.assign(ov, iof, 0.5, pos_iou_thr=0.5, neg_iou_thr=0.4)
    assert got.tolist() == [1, -1, -1, -1] and mx.tolist() == [np.float32(0.7), -1.0, -1.0, np.float32(0.45)]
    got, _ = A.assign(ov, iof, -1, pos_iou_thr=0.5, neg_iou_thr=0.4)           # disabled (the shipped configs)
    assert got.tolist() == [1, 1, 0, -1]
