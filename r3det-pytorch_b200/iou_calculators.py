"""mmdet-registered IoU calculators RBboxOverlaps2D_v1/v2/v3 — mirror of
r3det/core/bbox/iou_calculators/rotate_iou2d_calculator.py (:7-48, :83-124, :159-199) and the
rbbox_overlaps_v1/v2/v3 functions (:51-80, :127-156, :202-231).

When mmdet is importable the classes are registered into its IOU_CALCULATORS registry under the
reference's names, so `iou_calculator=dict(type='RBboxOverlaps2D_v1')` in a config resolves to them;
otherwise a minimal local registry with the same `register_module` / `build` surface is used.
v2 replaces mmcv.ops.box_iou_rotated (third-party; restated from the in-tree ml_nms_rotated header,
parity unpinned upstream — SURVEY.md §8c)."""
import torch

from . import _lib as L
from .box_iou_rotated import obb_overlaps
from .rbbox_geo import aligned_iou, pairwise_iou, rbbox_iou

try:  # pragma: no cover - mmdet is not installed in the build image
    from mmdet.core.bbox.iou_calculators.builder import IOU_CALCULATORS
except Exception:  # noqa: BLE001
    class _Registry(dict):
        def register_module(self, name=None, force=False):
            def deco(cls):
                key = name or cls.__name__
                if key in self and not force:
                    raise KeyError(f'{key} is already registered')
                self[key] = cls
                return cls
            return deco

        def build(self, cfg):
            cfg = dict(cfg)
            return self[cfg.pop('type')](**cfg)

    IOU_CALCULATORS = _Registry()


def _strip(bboxes1, bboxes2):
    assert bboxes1.size(-1) in [0, 5, 6]
    assert bboxes2.size(-1) in [0, 5, 6]
    if bboxes2.size(-1) == 6:
        bboxes2 = bboxes2[..., :5]
    if bboxes1.size(-1) == 6:
        bboxes1 = bboxes1[..., :5]
    return bboxes1, bboxes2


def _check(bboxes1, bboxes2, mode, is_aligned):
    assert mode in ['iou', 'iof']
    assert (bboxes1.size(-1) == 5 or bboxes1.size(0) == 0)
    assert (bboxes2.size(-1) == 5 or bboxes2.size(0) == 0)
    rows, cols = bboxes1.size(0), bboxes2.size(0)
    if is_aligned:
        assert rows == cols
    return rows, cols


def rbbox_overlaps_v1(bboxes1, bboxes2, mode='iou', is_aligned=False):
    rows, cols = _check(bboxes1, bboxes2, mode, is_aligned)
    if rows * cols == 0:
        # the reference returns UNINITIALISED memory here (rotate_iou2d_calculator.py:77-78); zeros are a subset
        return bboxes1.new_zeros(rows, 1) if is_aligned else bboxes1.new_zeros(rows, cols)
    return rbbox_iou(bboxes1, bboxes2, is_aligned, (mode == 'iof'))


def rbbox_overlaps_v2(bboxes1, bboxes2, mode='iou', is_aligned=False):
    rows, cols = _check(bboxes1, bboxes2, mode, is_aligned)
    if rows * cols == 0:
        return bboxes1.new_zeros(rows, 1) if is_aligned else bboxes1.new_zeros(rows, cols)
    if is_aligned:
        return aligned_iou(bboxes1, bboxes2, 'v2', mode)
    return pairwise_iou(bboxes1, bboxes2, 'v2', mode)


def rbbox_overlaps_v3(bboxes1, bboxes2, mode='iou', is_aligned=False):
    rows, cols = _check(bboxes1, bboxes2, mode, is_aligned)
    if rows * cols == 0:
        return bboxes1.new_zeros(rows, 1) if is_aligned else bboxes1.new_zeros(rows, cols)
    return obb_overlaps(bboxes1, bboxes2, mode, is_aligned)


def _make(name, fn, contiguous):
    class _Calc(object):
        """2D Overlaps (IoU / IoF) calculator for rotated boxes <cx, cy, w, h, a[, score]>."""

        def __call__(self, bboxes1, bboxes2, mode='iou', is_aligned=False, version='v1'):
            bboxes1, bboxes2 = _strip(bboxes1, bboxes2)
            if contiguous:
                bboxes1, bboxes2 = bboxes1.contiguous(), bboxes2.contiguous()
            return fn(bboxes1, bboxes2, mode, is_aligned)

        def __repr__(self):
            return self.__class__.__name__ + '()'

    _Calc.__name__ = _Calc.__qualname__ = name
    return IOU_CALCULATORS.register_module(name=name, force=True)(_Calc)


RBboxOverlaps2D_v1 = _make('RBboxOverlaps2D_v1', rbbox_overlaps_v1, True)
RBboxOverlaps2D_v2 = _make('RBboxOverlaps2D_v2', rbbox_overlaps_v2, True)
RBboxOverlaps2D_v3 = _make('RBboxOverlaps2D_v3', rbbox_overlaps_v3, False)
