"""FRM on B200 — drop-in for r3det/ops/fr/feature_refine_module.py (names, signatures, state_dict keys).

    feature_refine(features, best_rbboxes, spatial_scale, points=1)       autograd op (:10-43)
    FR(spatial_scale, points=1)                                           nn.Module around it (:46-63)
    FeatureRefineModule(in_channels, featmap_strides, conv_cfg, norm_cfg) x + FR(conv_5_1(conv_1_5(x)) + conv_1_1(x)) (:66-127)

The op is the CUDA pair r3g_frm_forward_f32 / r3g_frm_backward_f32 (csrc/frm.cu): it writes every output element
itself (no zero-filled buffer is passed in, unlike :21-23 / :36-39), the backward is atomic-free and only the
features receive a gradient.  The three convolutions stay ordinary cuDNN nn.Conv2d layers."""
import ctypes as C

import torch
from torch import nn
from torch.autograd.function import Function, once_differentiable

from . import _lib as L


def _prep(x, boxes):
    L.require_cuda(x, boxes)
    if x.dim() != 4:
        raise ValueError('features must be (N, C, H, W)')
    xf = x.contiguous() if x.dtype == torch.float32 else x.float().contiguous()
    bf = boxes.float().contiguous()
    n, c, h, w = xf.shape
    if bf.numel() != n * h * w * 5:
        raise ValueError(f'best_rbboxes must hold one <x,y,w,h,a> per location: expected {n * h * w} boxes')
    return xf, bf, (n, c, h, w)


def frm_forward(features, best_rbboxes, spatial_scale, points=1):
    """out = features + sum over the sample points of bilinear(features, y_p, x_p)."""
    xf, bf, (n, c, h, w) = _prep(features, best_rbboxes)
    out = torch.empty_like(xf)
    with L.device_guard(xf.device):
        L.check(L.lib().r3g_frm_forward_f32(L.ptr(xf), L.ptr(bf), n, c, h, w, float(spatial_scale), int(points),
                                            L.ptr(out), L.stream_ptr(xf.device)))
    return out if features.dtype == torch.float32 else out.to(features.dtype)


def frm_backward(grad_output, best_rbboxes, spatial_scale, points=1):
    """grad wrt features of frm_forward (transpose of the same linear map)."""
    gf, bf, (n, c, h, w) = _prep(grad_output, best_rbboxes)
    gin = torch.empty_like(gf)
    lib = L.lib()
    need = C.c_size_t(0)
    L.check(lib.r3g_frm_backward_workspace_bytes(n, h, w, int(points), C.byref(need)))
    ws = L.workspace(need.value, gf.device)
    with L.device_guard(gf.device):
        L.check(lib.r3g_frm_backward_f32(L.ptr(gf), L.ptr(bf), n, c, h, w, float(spatial_scale), int(points),
                                         L.ptr(gin), L.ptr(ws), ws.numel(), L.stream_ptr(gf.device)))
    return gin if grad_output.dtype == torch.float32 else gin.to(grad_output.dtype)


class FeatureRefineFunction(Function):
    @staticmethod
    def forward(ctx, features, best_rbboxes, spatial_scale, points=1):
        assert points in [1, 5]
        assert features.is_cuda
        ctx.frm = (spatial_scale, points)
        ctx.save_for_backward(best_rbboxes)
        return frm_forward(features, best_rbboxes, spatial_scale, points)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        assert grad_output.is_cuda
        (boxes,) = ctx.saved_tensors
        scale, points = ctx.frm
        g = frm_backward(grad_output, boxes, scale, points) if ctx.needs_input_grad[0] else None
        return g, None, None, None


feature_refine = FeatureRefineFunction.apply


class FR(nn.Module):
    def __init__(self, spatial_scale, points=1):
        super().__init__()
        self.spatial_scale, self.points = float(spatial_scale), points

    def forward(self, features, best_rbboxes):
        return feature_refine(features, best_rbboxes, self.spatial_scale, self.points)

    def extra_repr(self):
        return f'spatial_scale={self.spatial_scale}, points={self.points}'

    def __repr__(self):
        return f'{type(self).__name__}({self.extra_repr()})'


class FeatureRefineModule(nn.Module):
    """Per FPN level: refined = x + FR(conv_5_1(conv_1_5(x)) + conv_1_1(x), boxes of that level)."""

    _CONVS = (('conv_5_1', (5, 1), (2, 0)), ('conv_1_5', (1, 5), (0, 2)), ('conv_1_1', 1, 0))

    def __init__(self, in_channels, featmap_strides, conv_cfg=None, norm_cfg=None):
        super().__init__()
        self.in_channels, self.featmap_strides = in_channels, featmap_strides
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        self.fr = nn.ModuleList(FR(spatial_scale=1 / s) for s in featmap_strides)
        for name, k, pad in self._CONVS:
            setattr(self, name, nn.Conv2d(in_channels, in_channels, kernel_size=k, stride=1, padding=pad))

    def init_weights(self):
        """what mmcv.cnn.normal_init(m, std=0.01) does: N(0, 0.01) weights, zero bias."""
        for name, _, _ in self._CONVS:
            conv = getattr(self, name)
            nn.init.normal_(conv.weight, mean=0.0, std=0.01)
            if conv.bias is not None:
                nn.init.zeros_(conv.bias)

    def forward(self, x, best_rbboxes):
        """x: per-level feature maps; best_rbboxes[img][lvl]: (H*W, 5) refined boxes of one image and level."""
        per_level_boxes = [torch.cat(level) for level in zip(*best_rbboxes)]
        refined = []
        for feat, boxes, fr in zip(x, per_level_boxes, self.fr):
            mixed = self.conv_5_1(self.conv_1_5(feat)) + self.conv_1_1(feat)
            refined.append(feat + fr(mixed, boxes))
        return refined
