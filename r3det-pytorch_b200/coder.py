"""Rotated box coder — mirror of r3det/core/bbox/coder/delta_xywha_rbbox_coder.py.

DeltaXYWHAOBBoxCoder(target_means, target_stds, angle_range, add_ctr_clamp, ctr_clamp) with .encode / .decode, and the
six module-level functions bbox2delta_v1/v2/v3, delta2bbox_v1/v2/v3 with the reference's signatures.  Each call is one
CUDA kernel (csrc/coder.cu) instead of ~25 small torch kernels.  CUDA tensors only, no CPU fallback."""
import ctypes as C

import torch

from . import _lib as L

_VER = {'v1': 1, 'v2': 2, 'v3': 3}
_ZERO5 = (0., 0., 0., 0., 0.)
_ONE5 = (1., 1., 1., 1., 1.)


def _f5(x):
    vals = [float(v) for v in x]
    assert len(vals) == 5
    return (C.c_float * 5)(*vals)


def _encode(proposals, gt, means, stds, version):
    if version not in _VER:
        raise NotImplementedError
    L.require_cuda(proposals, gt)
    if version == 'v1':
        assert proposals.size() == gt.size()                              # delta_xywha_rbbox_coder.py:123
    p, ps = L.as_f32_rows(proposals.reshape(-1, proposals.size(-1)))
    g, gs = L.as_f32_rows(gt.reshape(-1, gt.size(-1)))
    n = p.size(0)
    out = torch.empty((n, 5), dtype=torch.float32, device=p.device)
    if n:
        with L.device_guard(p.device):
            L.check(L.lib().r3g_bbox2delta_f32(L.ptr(p), ps, L.ptr(g), gs, n, _f5(means), _f5(stds), _VER[version], L.ptr(out),
                                               L.stream_ptr(p.device)))
    return out.reshape(*proposals.shape[:-1], 5)


def _decode(rois, deltas, means, stds, version, max_shape=None, wh_ratio_clip=16 / 1000, add_ctr_clamp=False, ctr_clamp=32):
    if version not in _VER:
        raise NotImplementedError
    L.require_cuda(rois, deltas)
    assert rois.dim() == 2 and deltas.dim() == 2 and deltas.size(1) % 5 == 0 and deltas.size(0) == rois.size(0)
    r, rs = L.as_f32_rows(rois)
    d, _ = L.as_f32_rows(deltas)
    n, groups = d.size(0), d.size(1) // 5
    out = torch.empty_like(d)
    ms = None
    if max_shape is not None and version == 'v1':                         # v2 / v3 ignore max_shape (:97-99)
        ms = (C.c_int * 2)(int(max_shape[0]), int(max_shape[1]))
    if n:
        with L.device_guard(r.device):
            L.check(L.lib().r3g_delta2bbox_f32(L.ptr(r), n, rs, L.ptr(d), groups, _f5(means), _f5(stds), _VER[version], ms,
                                               float(wh_ratio_clip), int(bool(add_ctr_clamp)), float(ctr_clamp), L.ptr(out),
                                               L.stream_ptr(r.device)))
    return out


def bbox2delta_v1(proposals, gt, means=_ZERO5, stds=_ONE5):
    """delta_xywha_rbbox_coder.py:104-139"""
    return _encode(proposals, gt, means, stds, 'v1')


def bbox2delta_v2(proposals, gt, means=_ZERO5, stds=_ONE5):
    """delta_xywha_rbbox_coder.py:214-252"""
    return _encode(proposals, gt, means, stds, 'v2')


def bbox2delta_v3(proposals, gt, means=_ZERO5, stds=_ONE5):
    """delta_xywha_rbbox_coder.py:314-360"""
    return _encode(proposals, gt, means, stds, 'v3')


def delta2bbox_v1(rois, deltas, means=_ZERO5, stds=_ONE5, max_shape=None, wh_ratio_clip=16 / 1000, add_ctr_clamp=False,
                  ctr_clamp=32):
    """delta_xywha_rbbox_coder.py:142-211"""
    return _decode(rois, deltas, means, stds, 'v1', max_shape, wh_ratio_clip, add_ctr_clamp, ctr_clamp)


def delta2bbox_v2(proposals, deltas, means=_ZERO5, stds=_ONE5, wh_ratio_clip=16 / 1000):
    """delta_xywha_rbbox_coder.py:255-311"""
    return _decode(proposals, deltas, means, stds, 'v2', None, wh_ratio_clip)


def delta2bbox_v3(proposals, deltas, means=_ZERO5, stds=_ONE5, wh_ratio_clip=16 / 1000):
    """delta_xywha_rbbox_coder.py:363-423"""
    return _decode(proposals, deltas, means, stds, 'v3', None, wh_ratio_clip)


class DeltaXYWHAOBBoxCoder:
    """Same constructor arguments and methods as the reference class (delta_xywha_rbbox_coder.py:11-101)."""

    def __init__(self, target_means=_ZERO5, target_stds=_ONE5, angle_range='v1', add_ctr_clamp=False, ctr_clamp=32):
        self.means = target_means
        self.stds = target_stds
        self.add_ctr_clamp = add_ctr_clamp
        self.ctr_clamp = ctr_clamp
        self.angle_range = angle_range

    def encode(self, bboxes, gt_bboxes):
        assert bboxes.size(0) == gt_bboxes.size(0)
        assert bboxes.size(-1) == 5
        assert gt_bboxes.size(-1) == 5
        return _encode(bboxes, gt_bboxes, self.means, self.stds, self.angle_range)

    def decode(self, bboxes, pred_bboxes, max_shape=None, wh_ratio_clip=16 / 1000):
        assert pred_bboxes.size(0) == bboxes.size(0)
        return _decode(bboxes, pred_bboxes, self.means, self.stds, self.angle_range, max_shape, wh_ratio_clip,
                       self.add_ctr_clamp, self.ctr_clamp)

    def c_args(self, wh_ratio_clip=16 / 1000):
        """(means, stds, variant, wh_ratio_clip, add_ctr_clamp, ctr_clamp) as the C ABI takes them."""
        if self.angle_range not in _VER:
            raise NotImplementedError
        return (_f5(self.means), _f5(self.stds), _VER[self.angle_range], float(wh_ratio_clip), int(bool(self.add_ctr_clamp)),
                float(self.ctr_clamp))


try:  # pragma: no cover - mmdet is not installed in the build image
    from mmdet.core.bbox.builder import BBOX_CODERS
    # same registry name as the reference class, so `bbox_coder=dict(type='DeltaXYWHAOBBoxCoder', ...)` in every config under
    # configs/** resolves to this implementation once r3det_b200 is imported after (or instead of) r3det.core.bbox.coder
    BBOX_CODERS.register_module(name='DeltaXYWHAOBBoxCoder', force=True)(DeltaXYWHAOBBoxCoder)
except Exception:  # noqa: BLE001
    pass
