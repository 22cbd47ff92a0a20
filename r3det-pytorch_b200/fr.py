"""FRM — host-side mirror of r3det/ops/fr/feature_refine_module.py:10-127.

FeatureRefineFunction / feature_refine / FR / FeatureRefineModule keep the reference's signatures:
forward(features (N,C,H,W), best_rbboxes (N*H*W,5), spatial_scale, points in {1,5}); backward is
`once_differentiable` and returns a gradient for the features only (:26-40).  The three convolutions of
FeatureRefineModule stay on cuDNN (out of scope); only the FR op is replaced."""
import ctypes as C

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L


def frm_forward(features, best_rbboxes, spatial_scale, points=1):
    L.require_cuda(features, best_rbboxes)
    assert features.dim() == 4
    feats = features.contiguous()
    if feats.dtype != torch.float32:
        feats = feats.float()
    boxes = best_rbboxes.float().contiguous()
    N, Cc, H, W = feats.shape
    assert boxes.numel() == N * H * W * 5, 'best_rbboxes must hold one 5-tuple per location'
    out = torch.empty_like(feats)
    with L.device_guard(feats.device):
        L.check(L.lib().r3g_frm_forward_f32(L.ptr(feats), L.ptr(boxes), N, Cc, H, W, float(spatial_scale), int(points),
                                            L.ptr(out), L.stream_ptr(feats.device)))
    return out.to(features.dtype)


def frm_backward(grad_output, best_rbboxes, spatial_scale, points=1):
    L.require_cuda(grad_output, best_rbboxes)
    g = grad_output.contiguous()
    if g.dtype != torch.float32:
        g = g.float()
    boxes = best_rbboxes.float().contiguous()
    N, Cc, H, W = g.shape
    gin = torch.empty_like(g)
    lib = L.lib()
    nbytes = C.c_size_t(0)
    L.check(lib.r3g_frm_backward_workspace_bytes(N, H, W, int(points), C.byref(nbytes)))
    ws = L.workspace(nbytes.value, g.device)
    with L.device_guard(g.device):
        L.check(lib.r3g_frm_backward_f32(L.ptr(g), L.ptr(boxes), N, Cc, H, W, float(spatial_scale), int(points),
                                         L.ptr(gin), L.ptr(ws), ws.numel(), L.stream_ptr(g.device)))
    return gin.to(grad_output.dtype)


class FeatureRefineFunction(Function):
    """Feature refine class (feature_refine_module.py:10-40)."""

    @staticmethod
    def forward(ctx, features, best_rbboxes, spatial_scale, points=1):
        ctx.spatial_scale = spatial_scale
        ctx.points = points
        ctx.save_for_backward(best_rbboxes)
        assert points in [1, 5]
        assert features.is_cuda
        return frm_forward(features, best_rbboxes, spatial_scale, points)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        best_rbboxes = ctx.saved_tensors[0]
        assert grad_output.is_cuda
        grad_input = None
        if ctx.needs_input_grad[0]:
            grad_input = frm_backward(grad_output, best_rbboxes, ctx.spatial_scale, ctx.points)
        return grad_input, None, None, None


feature_refine = FeatureRefineFunction.apply


class FR(nn.Module):
    """FR module (feature_refine_module.py:46-63)."""

    def __init__(self, spatial_scale, points=1):
        super(FR, self).__init__()
        self.spatial_scale = float(spatial_scale)
        self.points = points

    def forward(self, features, best_rbboxes):
        return feature_refine(features, best_rbboxes, self.spatial_scale, self.points)

    def __repr__(self):
        return f'{self.__class__.__name__}(spatial_scale={self.spatial_scale}, points={self.points})'


class FeatureRefineModule(nn.Module):
    """Feature refine module (feature_refine_module.py:66-127): x + FR(conv_5_1(conv_1_5(x)) + conv_1_1(x), boxes)."""

    def __init__(self, in_channels, featmap_strides, conv_cfg=None, norm_cfg=None):
        super(FeatureRefineModule, self).__init__()
        self.in_channels = in_channels
        self.featmap_strides = featmap_strides
        self.conv_cfg = conv_cfg
        self.norm_cfg = norm_cfg
        self._init_layers()

    def _init_layers(self):
        self.fr = nn.ModuleList([FR(spatial_scale=1 / s) for s in self.featmap_strides])
        self.conv_5_1 = nn.Conv2d(self.in_channels, self.in_channels, kernel_size=(5, 1), stride=1, padding=(2, 0))
        self.conv_1_5 = nn.Conv2d(self.in_channels, self.in_channels, kernel_size=(1, 5), stride=1, padding=(0, 2))
        self.conv_1_1 = nn.Conv2d(self.in_channels, self.in_channels, kernel_size=1)

    def init_weights(self):
        """mmcv.cnn.normal_init(m, std=0.01): N(0, 0.01) weights, zero bias."""
        for m in (self.conv_5_1, self.conv_1_5, self.conv_1_1):
            nn.init.normal_(m.weight, 0, 0.01)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward(self, x, best_rbboxes):
        """x: list of per-level feature maps; best_rbboxes: list (image) of list (level) of (H*W, 5) boxes."""
        mlvl_rbboxes = [torch.cat(best_rbbox) for best_rbbox in zip(*best_rbboxes)]
        out = []
        for x_scale, best_rbboxes_scale, fr_scale in zip(x, mlvl_rbboxes, self.fr):
            feat_scale_1 = self.conv_5_1(self.conv_1_5(x_scale))
            feat_scale_2 = self.conv_1_1(x_scale)
            feat_scale = feat_scale_1 + feat_scale_2
            feat_refined_scale = fr_scale(feat_scale, best_rbboxes_scale)
            out.append(x_scale + feat_refined_scale)
        return out
