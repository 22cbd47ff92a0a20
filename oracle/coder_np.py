"""TEST INFRASTRUCTURE ONLY — float32 numpy restatement of the reference's rotated box coder and of the per-level
selection + decode that precedes multiclass NMS in `get_bboxes`:

  DeltaXYWHAOBBoxCoder.encode / .decode      r3det/core/bbox/coder/delta_xywha_rbbox_coder.py:46-101
    bbox2delta_v1 :104-139   delta2bbox_v1 :142-211   bbox2delta_v2 :214-252   delta2bbox_v2 :255-311
    bbox2delta_v3 :314-360   delta2bbox_v3 :363-423
  RAnchorHead._get_bboxes_single (up to the NMS call)   r3det/models/dense_heads/rotate_anchor_head.py:590-673
  RRetinaHead.filter_bboxes                  r3det/models/dense_heads/rotate_retina_head.py:117-179
  RRetinaRefineHead.refine_bboxes            r3det/models/dense_heads/rotate_retina_refine_head.py:56-97

Every arithmetic step is rounded to float32 in the order the torch expressions evaluate.  Pinned against the reference
functions themselves through tests/golden/coder_ref.npz (tests/golden/make_golden_coder.py).  Never imported by the
product."""
import numpy as np

F = np.float32
PI = F(np.pi)
HPI = F(np.pi / 2)
QPI = F(np.pi / 4)


def _rem(a, b):
    """torch.remainder (sign of the divisor) in float32."""
    a = np.asarray(a, F)
    m = np.fmod(a, F(b)).astype(F)
    fix = (m != 0) & ((F(b) < 0) != (m < 0))
    return np.where(fix, m + F(b), m).astype(F)


def _norm(deltas, means, stds):
    return ((deltas - np.asarray(means, F)[None]) / np.asarray(stds, F)[None]).astype(F)


def bbox2delta(proposals, gt, means=(0.,) * 5, stds=(1.,) * 5, variant='v1'):
    p = np.asarray(proposals, F); g = np.asarray(gt, F)
    px, py, pw, ph, pa = (p[:, i] for i in range(5))
    gx, gy, gw, gh, ga = (g[:, i] for i in range(5))
    if variant == 'v1':                                                   # :127-139
        d = np.stack([(gx - px) / pw, (gy - py) / ph, np.log(gw / pw).astype(F), np.log(gh / ph).astype(F), ga - pa], -1)
        return _norm(d.astype(F), means, stds)
    if variant == 'v2':                                                   # :232-252
        c, s = np.cos(pa).astype(F), np.sin(pa).astype(F)
        dx = (c * (gx - px) + s * (gy - py)) / pw
        dy = (-s * (gx - px) + c * (gy - py)) / ph
        dt = _rem((ga - pa) + QPI, PI) - QPI
        dt = dt / PI
        d = np.stack([dx, dy, np.log(gw / pw).astype(F), np.log(gh / ph).astype(F), dt], -1)
        return _norm(d.astype(F), means, stds)
    # v3 :333-360
    d1 = _rem((ga - pa) + HPI, PI) - HPI
    d2 = _rem(((ga - pa) + HPI) + HPI, PI) - HPI
    first = np.abs(d1) < np.abs(d2)
    gw_r = np.where(first, gw, gh); gh_r = np.where(first, gh, gw); dt = np.where(first, d1, d2)
    c, s = np.cos(-pa).astype(F), np.sin(-pa).astype(F)
    dx = (c * (gx - px) + s * (gy - py)) / pw
    dy = (-s * (gx - px) + c * (gy - py)) / ph
    d = np.stack([dx, dy, np.log(gw_r / pw).astype(F), np.log(gh_r / ph).astype(F), dt], -1)
    return _norm(d.astype(F), means, stds)


def delta2bbox(rois, deltas, means=(0.,) * 5, stds=(1.,) * 5, variant='v1', max_shape=None, wh_ratio_clip=16 / 1000,
               add_ctr_clamp=False, ctr_clamp=32):
    r = np.asarray(rois, F); d = np.asarray(deltas, F)
    n, k = d.shape[0], d.shape[1] // 5
    d = d.reshape(n, k, 5)
    dn = (d * np.asarray(stds, F)[None, None] + np.asarray(means, F)[None, None]).astype(F)
    dx, dy, dw, dh, da = (dn[..., i] for i in range(5))
    px, py, pw, ph, pa = (r[:, i][:, None] for i in range(5))
    max_ratio = F(np.abs(np.log(wh_ratio_clip)))
    if variant == 'v1':                                                   # :171-211
        dxw = (pw * dx).astype(F); dyh = (ph * dy).astype(F)
        if add_ctr_clamp:
            dxw = np.clip(dxw, -F(ctr_clamp), F(ctr_clamp)); dyh = np.clip(dyh, -F(ctr_clamp), F(ctr_clamp))
            dw = np.minimum(dw, max_ratio); dh = np.minimum(dh, max_ratio)
        else:
            dw = np.clip(dw, -max_ratio, max_ratio); dh = np.clip(dh, -max_ratio, max_ratio)
        gw = pw * np.exp(dw).astype(F); gh = ph * np.exp(dh).astype(F)
        gx = px + dxw; gy = py + dyh; ga = pa + da
        if max_shape is not None:
            gx = np.clip(gx, F(0), F(max_shape[1] - 1)); gy = np.clip(gy, F(0), F(max_shape[0] - 1))
        return np.stack([gx, gy, gw, gh, ga], -1).astype(F).reshape(n, k * 5)
    dw = np.clip(dw, -max_ratio, max_ratio); dh = np.clip(dh, -max_ratio, max_ratio)
    if variant == 'v2':                                                   # :283-311
        dt = (da * PI).astype(F)
        c, s = np.cos(pa).astype(F), np.sin(pa).astype(F)
        gx = ((dx * pw).astype(F) * c - (dy * ph).astype(F) * s).astype(F) + px
        gy = ((dx * pw).astype(F) * s + (dy * ph).astype(F) * c).astype(F) + py
        gw = pw * np.exp(dw).astype(F); gh = ph * np.exp(dh).astype(F)
        gt = _rem((dt + pa) + QPI, PI) - QPI
        return np.stack([gx, gy, gw, gh, gt], -1).astype(F).reshape(n, k * 5)
    # v3 :391-423
    c, s = np.cos(-pa).astype(F), np.sin(-pa).astype(F)
    gx = ((dx * pw).astype(F) * c - (dy * ph).astype(F) * s).astype(F) + px
    gy = ((dx * pw).astype(F) * s + (dy * ph).astype(F) * c).astype(F) + py
    gw = (pw * np.exp(dw).astype(F)).astype(F); gh = (ph * np.exp(dh).astype(F)).astype(F)
    gt = (da + pa).astype(F)
    big = gw > gh
    w_r = np.where(big, gw, gh); h_r = np.where(big, gh, gw)
    t_r = np.where(big, gt, gt + HPI).astype(F)
    t_r = _rem(t_r + HPI, PI) - HPI
    return np.stack([gx, gy, w_r, h_r, t_r], -1).astype(F).reshape(n, k * 5)


def sigmoid(x):
    x = np.asarray(x, F)
    return (F(1) / (F(1) + np.exp(-x).astype(F))).astype(F)


def rows_of_level(t, width):
    """(A*width, H, W) -> (H*W*A, width): `t.permute(1, 2, 0).reshape(-1, width)`."""
    return np.ascontiguousarray(np.transpose(np.asarray(t, F), (1, 2, 0))).reshape(-1, width)


def select_decode(cls_score_list, bbox_pred_list, mlvl_anchors, img_shape, scale_factor, nms_pre, num_classes, coder,
                  rescale=False):
    """rotate_anchor_head.py:626-662 for use_sigmoid_cls=True: returns (mlvl_bboxes (K,5), mlvl_scores (K, C+1)).
    `coder` = dict(means, stds, variant, add_ctr_clamp, ctr_clamp).  Top-k ties resolve to the lower row index."""
    boxes, scores = [], []
    for cs, bp, an in zip(cls_score_list, bbox_pred_list, mlvl_anchors):
        sc = sigmoid(rows_of_level(cs, num_classes))
        bp = rows_of_level(bp, 5)
        an = np.asarray(an, F)
        if nms_pre > 0 and sc.shape[0] > nms_pre:
            top = np.argsort(-sc.max(1), kind='stable')[:nms_pre]
            an, bp, sc = an[top], bp[top], sc[top]
        boxes.append(delta2bbox(an, bp, coder['means'], coder['stds'], coder['variant'], img_shape, 16 / 1000,
                                coder.get('add_ctr_clamp', False), coder.get('ctr_clamp', 32)))
        scores.append(sc)
    boxes = np.concatenate(boxes).astype(F)
    if rescale:
        boxes[:, :4] = boxes[:, :4] / np.asarray(scale_factor, F)[None]
    scores = np.concatenate(scores)
    return boxes, np.concatenate([scores, np.zeros((scores.shape[0], 1), F)], 1)


def filter_bboxes(cls_score, bbox_pred, anchors, num_anchors, num_classes, coder):
    """rotate_retina_head.py:148-177 for one level and one image: (A*C,H,W), (A*5,H,W), (H*W*A,5) -> (H*W, 5)."""
    cs = rows_of_level(cls_score, num_classes).reshape(-1, num_anchors, num_classes).max(-1)
    best = cs.argmax(-1)                                                  # first maximum
    bp = rows_of_level(bbox_pred, 5).reshape(-1, num_anchors, 5)
    an = np.asarray(anchors, F).reshape(-1, num_anchors, 5)
    idx = np.arange(len(best))
    return delta2bbox(an[idx, best], bp[idx, best], coder['means'], coder['stds'], coder['variant'], None, 16 / 1000,
                      coder.get('add_ctr_clamp', False), coder.get('ctr_clamp', 32))


def refine_bboxes(bbox_pred, rois, coder):
    """rotate_retina_refine_head.py:86-96 for one level and one image: (5,H,W), (H*W,5) -> (H*W,5)."""
    return delta2bbox(rois, rows_of_level(bbox_pred, 5), coder['means'], coder['stds'], coder['variant'], None, 16 / 1000,
                      coder.get('add_ctr_clamp', False), coder.get('ctr_clamp', 32))
