"""Rotated box coder — mirror of r3det/core/bbox/coder/delta_xywha_rbbox_coder.py.

DeltaXYWHAOBBoxCoder(target_means, target_stds, angle_range, add_ctr_clamp, ctr_clamp) with .encode / .decode, and the
six module-level functions bbox2delta_v1/v2/v3, delta2bbox_v1/v2/v3 with the reference's signatures.  Each call is one
CUDA kernel (csrc/coder.cu) instead of ~25 small torch kernels.  CUDA tensors only, no CPU fallback.

Autograd: the reference decode is differentiable torch code and rotate_anchor_head.py:418-420 back-propagates through
`bbox_coder.decode` when `reg_decoded_bbox=True`.  decode is therefore a torch.autograd.Function here: the forward is the CUDA
kernel, the backward the analytic gradient with respect to the deltas (the boxes being decoded — anchors / rois — get none,
and raise if they ask for one).  encode produces regression TARGETS; an input that requires grad raises instead of
returning a silently detached result."""
import ctypes as C
import math

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
    L.require_no_grad('coder.encode', proposals, gt)
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
    if torch.is_grad_enabled() and (deltas.requires_grad or rois.requires_grad):
        L.require_no_grad('coder.decode (rois)', rois)
        return _DecodeFn.apply(rois, deltas, tuple(float(v) for v in means), tuple(float(v) for v in stds), version,
                               None if max_shape is None else (int(max_shape[0]), int(max_shape[1])), float(wh_ratio_clip),
                               bool(add_ctr_clamp), float(ctr_clamp))
    return _decode_raw(rois, deltas, means, stds, version, max_shape, wh_ratio_clip, add_ctr_clamp, ctr_clamp)


def _decode_raw(rois, deltas, means, stds, version, max_shape=None, wh_ratio_clip=16 / 1000, add_ctr_clamp=False, ctr_clamp=32):
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


class _DecodeFn(torch.autograd.Function):
    """delta2bbox_v1/v2/v3 with the gradient of the reference's torch expressions with respect to `deltas`
    (delta_xywha_rbbox_coder.py:172-211, 283-311, 391-423; torch.clamp passes the gradient inside its closed range)."""

    @staticmethod
    def forward(ctx, rois, deltas, means, stds, version, max_shape, wh_ratio_clip, add_ctr_clamp, ctr_clamp):
        out = _decode_raw(rois, deltas.detach(), means, stds, version, max_shape, wh_ratio_clip, add_ctr_clamp, ctr_clamp)
        ctx.save_for_backward(rois.detach(), deltas.detach())
        ctx.cfg = (means, stds, version, max_shape, wh_ratio_clip, add_ctr_clamp, ctr_clamp)
        ctx.in_dtype = deltas.dtype
        return out

    @staticmethod
    def backward(ctx, grad):
        rois, deltas = ctx.saved_tensors
        means, stds, version, max_shape, wh_ratio_clip, add_ctr_clamp, ctr_clamp = ctx.cfg
        n, g5 = deltas.shape
        d = deltas.float().reshape(n, g5 // 5, 5)
        g = grad.float().reshape(n, g5 // 5, 5)
        std = d.new_tensor(stds)
        den = d * std + d.new_tensor(means)
        r = rois.float()
        pw, ph, pa = r[:, 2, None], r[:, 3, None], r[:, 4, None]
        max_ratio = abs(math.log(wh_ratio_clip))
        dx, dy, dw, dh = den[..., 0], den[..., 1], den[..., 2], den[..., 3]
        if version == 'v1' and add_ctr_clamp:                     # one-sided size clamp, centre shift clamped (:188-192)
            mw, mh = dw <= max_ratio, dh <= max_ratio
        else:
            mw, mh = dw.abs() <= max_ratio, dh.abs() <= max_ratio
        gw = pw * dw.clamp(min=None if (version == 'v1' and add_ctr_clamp) else -max_ratio, max=max_ratio).exp()
        gh = ph * dh.clamp(min=None if (version == 'v1' and add_ctr_clamp) else -max_ratio, max=max_ratio).exp()
        out = torch.empty_like(d)
        if version == 'v1':
            sx, sy = pw * dx, ph * dy
            mx, my = torch.ones_like(mw), torch.ones_like(mh)
            if add_ctr_clamp:
                mx, my = sx.abs() <= ctr_clamp, sy.abs() <= ctr_clamp
                sx, sy = sx.clamp(-ctr_clamp, ctr_clamp), sy.clamp(-ctr_clamp, ctr_clamp)
            if max_shape is not None:
                gx, gy = r[:, 0, None] + sx, r[:, 1, None] + sy
                mx = mx & (gx >= 0) & (gx <= max_shape[1] - 1)
                my = my & (gy >= 0) & (gy <= max_shape[0] - 1)
            out[..., 0] = g[..., 0] * pw * mx
            out[..., 1] = g[..., 1] * ph * my
            out[..., 2] = g[..., 2] * gw * mw
            out[..., 3] = g[..., 3] * gh * mh
            out[..., 4] = g[..., 4]
        else:
            ang = pa if version == 'v2' else -pa
            c, s_ = torch.cos(ang), torch.sin(ang)
            out[..., 0] = (g[..., 0] * c + g[..., 1] * s_) * pw
            out[..., 1] = (g[..., 1] * c - g[..., 0] * s_) * ph
            if version == 'v2':
                ggw, ggh = g[..., 2], g[..., 3]
                out[..., 4] = g[..., 4] * math.pi
            else:                                                  # v3 returns (long edge, short edge): route the gradients back (:414-416)
                swap = gw > gh
                ggw = torch.where(swap, g[..., 2], g[..., 3])
                ggh = torch.where(swap, g[..., 3], g[..., 2])
                out[..., 4] = g[..., 4]
            out[..., 2] = ggw * gw * mw
            out[..., 3] = ggh * gh * mh
        return None, (out * std).reshape(n, g5).to(ctx.in_dtype), None, None, None, None, None, None, None


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
