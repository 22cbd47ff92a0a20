"""FRM on B200 — drop-in for r3det/ops/fr/feature_refine_module.py (names, signatures, state_dict keys).

    feature_refine(features, best_rbboxes, spatial_scale, points=1)       autograd op (:10-43)
    feature_refine_multi(features[], best_rbboxes[], scales[], points, residuals[])   every FPN level in one launch
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


def _prep_levels(xs, boxes):
    if len(xs) != len(boxes) or not xs:
        raise ValueError('need one box tensor per level')
    if len(xs) > 8:
        raise ValueError('at most 8 levels per call')
    prepped = [_prep(x, b) for x, b in zip(xs, boxes)]
    n, c = prepped[0][2][0], prepped[0][2][1]
    for _, _, (n_, c_, _, _) in prepped:
        if (n_, c_) != (n, c):
            raise ValueError('all levels must share batch size and channel count')
    return prepped, n, c


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() if t is not None and t.numel() else 0 for t in tensors])


def frm_forward_multi(features, best_rbboxes, spatial_scales, points=1, residuals=None):
    """All levels in one launch: out_l = [residuals_l +] features_l + sum_p bilinear(features_l, y_p, x_p)."""
    prepped, n, c = _prep_levels(features, best_rbboxes)
    xs = [p[0] for p in prepped]; bs = [p[1] for p in prepped]
    outs = [torch.empty_like(x) for x in xs]
    rs = None
    if residuals is not None:
        rs = [r.contiguous() if r.dtype == torch.float32 else r.float().contiguous() for r in residuals]
        assert all(r.shape == x.shape and r.is_cuda for r, x in zip(rs, xs))
    nl = len(xs)
    hw = (C.c_int * (2 * nl))(*[v for x in xs for v in (x.size(2), x.size(3))])
    sc = (C.c_float * nl)(*[float(s) for s in spatial_scales])
    dev = xs[0].device
    with L.device_guard(dev):
        L.check(L.lib().r3g_frm_forward_multi_f32(nl, _ptr_array(xs), _ptr_array(bs), None if rs is None else _ptr_array(rs), n, c, hw,
                                                  sc, int(points), _ptr_array(outs), L.stream_ptr(dev)))
    return [o if f.dtype == torch.float32 else o.to(f.dtype) for o, f in zip(outs, features)]


def frm_backward_multi(grad_outputs, best_rbboxes, spatial_scales, points=1):
    """grad wrt the features of frm_forward_multi: one tap sort / CSR build and one gather launch for all levels."""
    prepped, n, c = _prep_levels(grad_outputs, best_rbboxes)
    gs = [p[0] for p in prepped]; bs = [p[1] for p in prepped]
    gins = [torch.empty_like(g) for g in gs]
    nl = len(gs)
    hw = (C.c_int * (2 * nl))(*[v for g in gs for v in (g.size(2), g.size(3))])
    sc = (C.c_float * nl)(*[float(s) for s in spatial_scales])
    lib = L.lib()
    need = C.c_size_t(0)
    L.check(lib.r3g_frm_backward_multi_workspace_bytes(nl, n, hw, int(points), C.byref(need)))
    dev = gs[0].device
    ws = L.workspace(need.value, dev)
    with L.device_guard(dev):
        L.check(lib.r3g_frm_backward_multi_f32(nl, _ptr_array(gs), _ptr_array(bs), n, c, hw, sc, int(points), _ptr_array(gins),
                                               L.ptr(ws), ws.numel(), L.stream_ptr(dev)))
    return [g if go.dtype == torch.float32 else g.to(go.dtype) for g, go in zip(gins, grad_outputs)]


class FrmBackwardPlan:
    """The box-only half of the backward (tap sort + per-target CSR) for a list of levels, built on a SIDE stream so that it
    overlaps the forward kernels; `apply(grads)` runs the gather on the current stream after waiting for the plan."""

    _side = {}

    def __init__(self, shapes, best_rbboxes, spatial_scales, points):
        self.points = int(points)
        self.n, self.c = shapes[0][0], shapes[0][1]
        self.boxes = [b.float().contiguous() for b in best_rbboxes]
        nl = len(shapes)
        self.nl = nl
        self.hw = (C.c_int * (2 * nl))(*[v for sh in shapes for v in (sh[2], sh[3])])
        self.sc = (C.c_float * nl)(*[float(s) for s in spatial_scales])
        dev = self.boxes[0].device
        self.dev = dev
        lib = L.lib()
        need = C.c_size_t(0)
        L.check(lib.r3g_frm_backward_multi_workspace_bytes(nl, self.n, self.hw, self.points, C.byref(need)))
        self.ws = L.workspace(need.value, dev)
        cur = torch.cuda.current_stream(dev)
        if torch.cuda.is_current_stream_capturing():
            # under CUDA-graph capture a forward whose backward never runs would leave side-stream work unjoined: the plan
            # is captured on the capturing stream itself (it is ~50 us; the overlap is given up, correctness is not)
            with L.device_guard(dev):
                L.check(lib.r3g_frm_backward_plan_multi_f32(nl, _ptr_array(self.boxes), self.n, self.hw, self.sc, self.points,
                                                            L.ptr(self.ws), self.ws.numel(), C.c_void_p(cur.cuda_stream)))
            self.ready = None
            return
        side = FrmBackwardPlan._side.get(dev.index)
        if side is None:
            side = FrmBackwardPlan._side[dev.index] = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)                                  # the boxes are produced on the current stream
        with L.device_guard(dev), torch.cuda.stream(side):
            L.check(lib.r3g_frm_backward_plan_multi_f32(nl, _ptr_array(self.boxes), self.n, self.hw, self.sc, self.points,
                                                        L.ptr(self.ws), self.ws.numel(), C.c_void_p(side.cuda_stream)))
            self.ready = side.record_event()
        for t in self.boxes + [self.ws]:
            t.record_stream(side)

    def apply(self, grad_outputs):
        gs = [g.contiguous() if g.dtype == torch.float32 else g.float().contiguous() for g in grad_outputs]
        gins = [torch.empty_like(g) for g in gs]
        cur = torch.cuda.current_stream(self.dev)
        if self.ready is not None:
            cur.wait_event(self.ready)
        with L.device_guard(self.dev):                         # the apply only READS the plan: it can run any number of times
            L.check(L.lib().r3g_frm_backward_apply_multi_f32(self.nl, _ptr_array(gs), _ptr_array(self.boxes), self.n, gs[0].size(1),
                                                             self.hw, self.sc, self.points, _ptr_array(gins), L.ptr(self.ws),
                                                             self.ws.numel(), L.stream_ptr(self.dev)))
        return [g if go.dtype == torch.float32 else g.to(go.dtype) for g, go in zip(gins, grad_outputs)]


class FeatureRefineMultiFunction(Function):
    """feature_refine over every FPN level at once: apply(points, scales, with_residual, *features, *boxes[, *residuals])
    -> tuple of refined maps.  Gradients: features get the FRM transpose, residuals the identity, boxes none
    (as in the reference, feature_refine_module.py:27-43)."""

    @staticmethod
    def forward(ctx, points, scales, with_residual, *tensors):
        assert points in [1, 5]
        nl = len(scales)
        feats, boxes = tensors[:nl], tensors[nl:2 * nl]
        resid = tensors[2 * nl:3 * nl] if with_residual else None
        assert all(f.is_cuda for f in feats)
        ctx.frm = (points, tuple(scales), nl, with_residual)
        ctx.save_for_backward(*boxes)
        need = [ctx.needs_input_grad[3 + l] for l in range(nl)]
        ctx.plan = None
        if any(need) and all(f.numel() for f in feats):
            # the backward's CSR depends on the boxes only: build it now, on a side stream, under the forward kernels
            idx = [l for l in range(nl) if need[l]]
            ctx.plan = (idx, FrmBackwardPlan([tuple(feats[l].shape) for l in idx], [boxes[l] for l in idx],
                                             [scales[l] for l in idx], points))
        return tuple(frm_forward_multi(feats, boxes, scales, points, resid))

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        points, scales, nl, with_residual = ctx.frm
        boxes = ctx.saved_tensors
        grads = [g.contiguous() for g in grads]
        need = [ctx.needs_input_grad[3 + l] for l in range(nl)]
        gin = [None] * nl
        if any(need):
            if ctx.plan is not None:
                idx, plan = ctx.plan
                res = plan.apply([grads[l] for l in idx])
            else:
                idx = [l for l in range(nl) if need[l]]
                res = frm_backward_multi([grads[l] for l in idx], [boxes[l] for l in idx], [scales[l] for l in idx], points)
            for l, g in zip(idx, res):
                gin[l] = g
        gres = [grads[l] if ctx.needs_input_grad[3 + 2 * nl + l] else None for l in range(nl)] if with_residual else []
        return (None, None, None, *gin, *([None] * nl), *gres)


def feature_refine_multi(features, best_rbboxes, spatial_scales, points=1, residuals=None):
    """List-in / list-out wrapper of FeatureRefineMultiFunction."""
    args = list(features) + list(best_rbboxes) + (list(residuals) if residuals is not None else [])
    return list(FeatureRefineMultiFunction.apply(points, tuple(float(s) for s in spatial_scales), residuals is not None, *args))


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
        mixed = [self.conv_5_1(self.conv_1_5(feat)) + self.conv_1_1(feat) for feat in x]
        # every level through one FRM launch, with the `x_scale + feat_refined_scale` add (:126) in its epilogue
        return feature_refine_multi(mixed, per_level_boxes, [fr.spatial_scale for fr in self.fr],
                                    self.fr[0].points if len(self.fr) else 1, residuals=list(x))
