"""The steps of the dense heads on either side of multiclass NMS and FRM — mirrors of

  RAnchorHead.get_bboxes / ._get_bboxes_single     r3det/models/dense_heads/rotate_anchor_head.py:519-673
  RRetinaHead.filter_bboxes                        r3det/models/dense_heads/rotate_retina_head.py:117-179
  RRetinaRefineHead.refine_bboxes                  r3det/models/dense_heads/rotate_retina_refine_head.py:56-97

as free functions that take the head's tensors and its coder (r3det_b200.coder.DeltaXYWHAOBBoxCoder).  The network
outputs are consumed in NCHW as the convolutions produce them; a whole batch is handled by one launch sequence
(csrc/coder.cu).  CUDA tensors only, sigmoid classification only (use_sigmoid_cls=True, as every reference config has)."""
import ctypes as C

import torch

from . import _lib as L
from .bbox_nms_rotated import multiclass_nms_rotated_batch, multiclass_nms_rotated_padded

MAX_IMAGES = 64     # per r3g_select_decode_f32 call


def _nchw(t):
    L.require_cuda(t)
    assert t.dim() == 4
    return t.detach().float().contiguous()


def select_decode(cls_scores, bbox_preds, mlvl_anchors, coder, nms_pre=-1, img_shapes=None, scale_factors=None):
    """Everything `_get_bboxes_single` does before the NMS call, for the whole batch.

    cls_scores[l] (B, A*C, H, W), bbox_preds[l] (B, A*5, H, W); mlvl_anchors[l] is (H*W*A, 5) shared by the batch or
    (B, H*W*A, 5) per image (the refine head's `bboxes_as_anchors`).  img_shapes: per image (H, W[, 3]) or None (no
    clamp); scale_factors: per image 4 floats, or None (no rescale).
    Returns mlvl_bboxes (B, K, 5) and mlvl_scores (B, K, C + 1) with the zero background column."""
    nl = len(cls_scores)
    assert nl == len(bbox_preds) == len(mlvl_anchors) and nl >= 1
    cls = [_nchw(t) for t in cls_scores]
    reg = [_nchw(t) for t in bbox_preds]
    dev = cls[0].device
    B = cls[0].size(0)
    assert reg[0].size(1) % 5 == 0
    A = reg[0].size(1) // 5
    assert cls[0].size(1) % A == 0
    Cn = cls[0].size(1) // A
    anc, strides, hw = [], [], []
    for l in range(nl):
        assert cls[l].shape[-2:] == reg[l].shape[-2:] and cls[l].size(0) == B == reg[l].size(0)
        assert cls[l].size(1) == A * Cn and reg[l].size(1) == A * 5
        H, W = cls[l].shape[-2:]
        a = mlvl_anchors[l]
        L.require_cuda(a)
        a = a.detach().float().contiguous()
        assert a.size(-1) == 5 and a.size(-2) == H * W * A
        if a.dim() == 3:
            assert a.size(0) == B
        anc.append(a); strides.append(a.size(1) * 5 if a.dim() == 3 else 0); hw += [H, W]
    if B > MAX_IMAGES:
        parts = [select_decode([c[s:s + MAX_IMAGES] for c in cls], [r[s:s + MAX_IMAGES] for r in reg],
                               [a[s:s + MAX_IMAGES] if a.dim() == 3 else a for a in anc], coder, nms_pre,
                               None if img_shapes is None else img_shapes[s:s + MAX_IMAGES],
                               None if scale_factors is None else scale_factors[s:s + MAX_IMAGES])
                 for s in range(0, B, MAX_IMAGES)]
        return torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    lib = L.lib()
    hw_c = (C.c_int64 * (2 * nl))(*hw)
    rows, nbytes = C.c_int64(), C.c_size_t()
    L.check(lib.r3g_select_decode_sizes(nl, B, A, hw_c, int(nms_pre), C.byref(rows), C.byref(nbytes)))
    K = rows.value
    boxes = torch.empty((B, K, 5), dtype=torch.float32, device=dev)
    scores = torch.empty((B, K, Cn + 1), dtype=torch.float32, device=dev)
    if B == 0 or K == 0:
        return boxes, scores
    ws = L.workspace(nbytes.value, dev)
    pp = lambda ts: (C.c_void_p * nl)(*[t.data_ptr() for t in ts])
    ms = None
    if img_shapes is not None:
        assert len(img_shapes) == B
        ms = (C.c_int * (2 * B))(*[int(v) for s in img_shapes for v in (s[0], s[1])])
    sf = None
    if scale_factors is not None:
        assert len(scale_factors) == B
        sf = (C.c_float * (4 * B))(*[float(v) for s in scale_factors for v in list(s)[:4]])
    means, stds, ver, clip, actr, ctr = coder.c_args()
    with L.device_guard(dev):
        L.check(lib.r3g_select_decode_f32(nl, pp(cls), pp(reg), pp(anc), (C.c_int64 * nl)(*strides), hw_c, B, A, Cn, int(nms_pre),
                                          means, stds, ver, clip, actr, ctr, ms, sf, L.ptr(boxes), L.ptr(scores), L.ptr(ws),
                                          ws.numel(), L.stream_ptr(dev)))
    return boxes, scores


def get_bboxes(cls_scores, bbox_preds, mlvl_anchors, img_metas, cfg, coder, rescale=False, with_nms=True):
    """RAnchorHead.get_bboxes (rotate_anchor_head.py:519-588): list over images of (det_bboxes (n, 6), det_labels (n,))
    — or (mlvl_bboxes, mlvl_scores) per image when with_nms is False.  `cfg` needs nms_pre, score_thr, nms, max_per_img
    (attribute or key access); img_metas[i] needs 'img_shape' and, with rescale, 'scale_factor'."""
    get = (lambda k, d=None: cfg.get(k, d)) if hasattr(cfg, 'get') else (lambda k, d=None: getattr(cfg, k, d))
    shapes = [m['img_shape'] for m in img_metas]
    sfs = [m['scale_factor'] for m in img_metas] if rescale else None
    boxes, scores = select_decode(cls_scores, bbox_preds, mlvl_anchors, coder, get('nms_pre', -1), shapes, sfs)
    if with_nms:
        return multiclass_nms_rotated_batch(boxes, scores, get('score_thr'), get('nms'), get('max_per_img'))
    return [(boxes[i], scores[i]) for i in range(boxes.size(0))]


def get_bboxes_padded(cls_scores, bbox_preds, mlvl_anchors, img_metas, cfg, coder, rescale=False):
    """`get_bboxes` with fixed-size outputs: dets (B, max_per_img, 6), labels (B, max_per_img), counts (B,) — the first
    counts[b] rows of image b are what `get_bboxes` returns for it.  Nothing is read back to the host between the network
    outputs and the detections (select + decode, candidate extraction, class-offset scales, segmented NMS and truncation are
    device kernels of libr3geo), so the whole tail can be captured in a CUDA graph and replayed."""
    get = (lambda k, d=None: cfg.get(k, d)) if hasattr(cfg, 'get') else (lambda k, d=None: getattr(cfg, k, d))
    shapes = [m['img_shape'] for m in img_metas]
    sfs = [m['scale_factor'] for m in img_metas] if rescale else None
    boxes, scores = select_decode(cls_scores, bbox_preds, mlvl_anchors, coder, get('nms_pre', -1), shapes, sfs)
    return multiclass_nms_rotated_padded(boxes, scores, get('score_thr'), get('nms'), get('max_per_img'))


def filter_bboxes(cls_scores, bbox_preds, mlvl_anchors, coder, as_batch=False):
    """RRetinaHead.filter_bboxes (rotate_retina_head.py:117-179): per location the anchor with the best class logit,
    decoded.  Returns bboxes_list[img][lvl] of (H*W, 5) like the reference, or with as_batch=True the per-level
    (B, H*W, 5) tensors (the layout FeatureRefineModule consumes without the per-image torch.cat)."""
    nl = len(cls_scores)
    assert nl == len(bbox_preds) == len(mlvl_anchors)
    lib = L.lib()
    means, stds, ver, clip, actr, ctr = coder.c_args()
    levels = []
    for l in range(nl):
        cls, reg = _nchw(cls_scores[l]), _nchw(bbox_preds[l])
        B, H, W = cls.size(0), cls.size(2), cls.size(3)
        assert reg.size(0) == B and reg.shape[-2:] == cls.shape[-2:] and reg.size(1) % 5 == 0
        A = reg.size(1) // 5
        assert cls.size(1) % A == 0
        a = mlvl_anchors[l].detach().float().contiguous()
        L.require_cuda(a)
        assert a.shape == (H * W * A, 5)
        out = torch.empty((B, H * W, 5), dtype=torch.float32, device=cls.device)
        if out.numel():
            with L.device_guard(cls.device):
                L.check(lib.r3g_filter_bboxes_f32(L.ptr(cls), L.ptr(reg), L.ptr(a), B, A, cls.size(1) // A, H, W, means, stds, ver,
                                                  clip, actr, ctr, L.ptr(out), L.stream_ptr(cls.device)))
        levels.append(out)
    if as_batch:
        return levels
    return [[lv[i] for lv in levels] for i in range(levels[0].size(0))] if levels else []


def refine_bboxes(cls_scores, bbox_preds, rois, coder, as_batch=False):
    """RRetinaRefineHead.refine_bboxes (rotate_retina_refine_head.py:56-97).  rois: rois[img][lvl] of (H*W, 5) as the
    reference passes them, or a list of per-level (B, H*W, 5) tensors."""
    nl = len(bbox_preds)
    assert nl == len(cls_scores)
    assert rois is not None
    lib = L.lib()
    means, stds, ver, clip, actr, ctr = coder.c_args()
    levels = []
    for l in range(nl):
        reg = _nchw(bbox_preds[l])
        B, H, W = reg.size(0), reg.size(2), reg.size(3)
        assert reg.size(1) == 5 and cls_scores[l].size(0) == B
        if isinstance(rois[0], (list, tuple)):
            r = torch.stack([rois[i][l] for i in range(B)])
        else:
            r = rois[l]
        L.require_cuda(r)
        r = r.detach().float().contiguous()
        assert r.shape == (B, H * W, 5)
        out = torch.empty_like(r)
        if out.numel():
            with L.device_guard(reg.device):
                L.check(lib.r3g_refine_bboxes_f32(L.ptr(reg), L.ptr(r), B, H, W, means, stds, ver, clip, actr, ctr, L.ptr(out),
                                                  L.stream_ptr(reg.device)))
        levels.append(out)
    if as_batch:
        return levels
    return [[lv[i] for lv in levels] for i in range(levels[0].size(0))] if levels else []
