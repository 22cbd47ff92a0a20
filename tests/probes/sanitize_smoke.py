"""compute-sanitizer target: small invocations of every kernel (odd sizes, tails, empty inputs)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import r3det_b200 as R
from r3det_b200._nms_core import nms_device
from r3det_b200.fr import frm_forward, frm_backward
from tests.util import rand_obb, clustered
dev = torch.device('cuda:0')
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
for v in ('v1', 'v2', 'v3'):
    for m, n in ((1, 1), (7, 129), (65, 300), (130, 515), (3, 1024)):
        a, b = rand_obb(m, 1, v), rand_obb(n, 2, v, 4, 300)
        for mode in ('iou', 'iof'):
            R.pairwise_iou(t(a), t(b), v, mode)
        R.aligned_iou(t(a), t(b[:m]), v)
    c, s, l = clustered(700, 3, v)
    for K in (1, 63, 64, 65, 700):
        nms_device(t(c[:K]), t(s[:K]), 0.1, v)
        nms_device(t(c[:K]), t(s[:K]), 0.1, v, labels=t(l[:K]), class_offset=torch.tensor(1025.0, device=dev), order_index=True)
    sc = torch.rand(300, 16, device=dev) ** 4
    R.multiclass_nms_rotated(t(c[:300]), sc, 0.05, dict(type=v, iou_thr=0.1), 100)
rng = np.random.default_rng(0)
for (N, Cc, H, W, stride) in ((1, 3, 1, 1, 128), (2, 5, 9, 13, 16), (1, 33, 40, 7, 8), (2, 16, 32, 32, 8)):
    feat = t(rng.standard_normal((N, Cc, H, W)).astype(np.float32))
    boxes = np.concatenate([rng.uniform(-50, W * stride + 50, (N * H * W, 1)), rng.uniform(-50, H * stride + 50, (N * H * W, 1)),
                            rng.uniform(1, 8 * stride, (N * H * W, 2)), rng.uniform(-1.6, 0, (N * H * W, 1))], 1).astype(np.float32)
    for P in (1, 5):
        frm_forward(feat, t(boxes), 1.0 / stride, P); frm_backward(feat, t(boxes), 1.0 / stride, P)
for v in ('v1', 'v2', 'v3'):
    o = t(rand_obb(1001, 4, v)); p = R.obb2poly(o, v); R.poly2obb(p, v); R.obb2hbb(o, v); R.obb2xyxy(o, v); R.hbb2obb(R.obb2xyxy(o, 'v3'), v)
# round-1 additions: fused assigner (incl. tie-list overflow), batched NMS / multiclass batch, large-K NMS path, multi-level
# FRM with residual, coder + dense-head tail, polygon NMS
from r3det_b200.fr import frm_forward_multi, frm_backward_multi
gt, an = t(rand_obb(37, 5, 'v1')), t(rand_obb(2051, 6, 'v1'))
for aa in (True, False):
    R.max_iou_assign(gt, an, 0.5, 0.4, 0.0, True, aa, 'v1')
R.max_iou_assign(gt[:2], t(np.repeat(rand_obb(2, 5, 'v1'), 40000, axis=0)), 0.5, 0.4, 0.0, True, True, 'v1')
R.max_iou_assign_batched([gt, gt[:0], gt[:5]], an, 0.5, 0.4, 0.0, True, True, 'v1')
R.max_iou_assign_batched([gt[:3], gt[:65]], torch.stack([an[:1001], an[1001:2002]]), 0.5, 0.4, 0.0, True, True, 'v3')
c, s, l = clustered(1500, 9, 'v1')
bid = torch.arange(3, device=dev).repeat_interleave(500)
nms_device(t(c), t(s), 0.1, 'v1', labels=t(l), class_offset=torch.tensor([1025.0, 900.0, 1100.0], device=dev), order_index=True,
           batch_ids=bid, n_batches=3)
cb, sb, lb = clustered(17000, 10, 'v3')                      # above the small-K path
nms_device(t(cb), t(sb), 0.1, 'v3', labels=t(lb), drop_small=True)
R.multiclass_nms_rotated_batch(t(c[:900]).reshape(3, 300, 5), torch.rand(3, 300, 16, device=dev) ** 4, 0.05, dict(type='v1', iou_thr=0.1), 50)
fs = [t(rng.standard_normal((2, 7, H, W)).astype(np.float32)) for H, W in ((9, 13), (5, 6), (1, 1))]
bs = [t(np.concatenate([rng.uniform(-20, 120, (2 * f.size(2) * f.size(3), 2)), rng.uniform(1, 60, (2 * f.size(2) * f.size(3), 2)),
                        rng.uniform(-1.6, 0, (2 * f.size(2) * f.size(3), 1))], 1).astype(np.float32)) for f in fs]
for P in (1, 5):
    frm_forward_multi(fs, bs, [1 / 8, 1 / 16, 1 / 64], P, residuals=fs); frm_backward_multi(fs, bs, [1 / 8, 1 / 16, 1 / 64], P)
for v in ('v1', 'v2', 'v3'):
    coder = R.DeltaXYWHAOBBoxCoder((0.,) * 5, (0.5,) * 5, angle_range=v)
    rois = t(rand_obb(777, 11, v)); d = torch.randn(777, 15, device=dev) * 0.3
    coder.decode(rois, d, max_shape=(500, 600)); coder.encode(rois, t(rand_obb(777, 12, v)))
    A, Cn = 3, 5
    cls = [torch.randn(2, A * Cn, H, W, device=dev) for H, W in ((9, 11), (4, 5), (2, 3))]
    reg = [torch.randn(2, A * 5, H, W, device=dev) * 0.2 for H, W in ((9, 11), (4, 5), (2, 3))]
    anc = [t(rand_obb(H * W * A, 13, v)) for H, W in ((9, 11), (4, 5), (2, 3))]
    metas = [dict(img_shape=(100, 120, 3), scale_factor=np.ones(4, np.float32))] * 2
    R.get_bboxes(cls, reg, anc, metas, dict(nms_pre=50, score_thr=0.05, nms=dict(type=v, iou_thr=0.1), max_per_img=20), coder, rescale=True)
    fl = R.filter_bboxes(cls, reg, anc, coder, as_batch=True)
    R.refine_bboxes(cls, [r[:, :5].contiguous() for r in reg], fl, coder)
from oracle import transforms_np as T
q = np.concatenate([T.obb2poly(c, 'v1') + rng.normal(0, 2, (1500, 8)).astype(np.float32), s[:, None]], 1).astype(np.float32)
for K in (1, 64, 65, 1500):
    R.poly_nms(t(q[:K]), 0.1)
R.poly_nms(t(q[:300]), 0.0)
# round-2 additions: NMS rounds kernel with several rounds per segment (K above the chunk size), polygon NMS on the rounds
# kernel, padded keep records / padded multiclass NMS / padded dense tail (device-side counts), TMA FRM (W % 4 == 0 levels,
# windows crossing every map edge) forward + backward
cb2, sb2, lb2 = clustered(6000, 21, 'v1')
bid2 = torch.arange(2, device=dev).repeat_interleave(3000)
kp, nm = nms_device(t(cb2), t(sb2), 0.1, 'v1', labels=t(lb2), class_offset=torch.tensor([1025.0, 900.0], device=dev), order_index=True,
                    batch_ids=bid2, n_batches=2)
from r3det_b200 import pack_keep_records
pack_keep_records(t(cb2), t(sb2), t(lb2), kp, nm, bid2, 2, 100)
# invalid image ids / labels of every kind, interleaved (they share one reserved segment key), on both ordering paths
for Kx in (900, 6000):
    lbx, bdx = lb2[:Kx].copy(), (np.arange(Kx) % 2).astype(np.int64)
    lbx[5::28] += 65536; bdx[12::28] += 65536; bdx[19::28] = -1; bdx[26::28] = 2
    nms_device(t(cb2[:Kx]), t(sb2[:Kx]), 0.1, 'v1', labels=t(lbx), class_offset=torch.tensor([1025.0, 900.0], device=dev), order_index=True,
               batch_ids=t(bdx), n_batches=2)
    nms_device(t(cb2[:Kx]), t(sb2[:Kx]), 0.1, 'v1', labels=t(lbx), class_offset=torch.tensor([1025.0, 900.0], device=dev), order_index=True,
               batch_ids=t(bdx), n_batches=2, sort_path=True)
# a broken R3G_NMS_LABEL_BITS promise (labels up to 14 with 2 bits): unspecified result, but no access outside the workspace
nms_device(t(cb2), t(sb2), 0.1, 'v1', labels=t(lb2), class_offset=torch.tensor(1025.0, device=dev), order_index=True, label_bits=2, sort_path=True)
nms_device(t(cb2), t(sb2), 0.3, 'v3')                         # one segment, K > chunk: several rounds
R.poly_nms(t(np.concatenate([T.obb2poly(cb2, 'v1'), sb2[:, None]], 1).astype(np.float32)), 0.1)
for v in ('v1', 'v2', 'v3'):
    R.multiclass_nms_rotated_padded(t(c[:900]).reshape(3, 300, 5), torch.rand(3, 300, 16, device=dev) ** 4, 0.05, dict(type=v, iou_thr=0.1), 50)
    coder = R.DeltaXYWHAOBBoxCoder((0.,) * 5, (0.5,) * 5, angle_range=v)
    R.get_bboxes_padded(cls, reg, anc, metas, dict(nms_pre=50, score_thr=0.05, nms=dict(type=v, iou_thr=0.1), max_per_img=20), coder, rescale=True)
for (N, Cc, H, W, stride) in ((2, 32, 12, 16, 8), (1, 64, 40, 44, 8), (2, 16, 4, 4, 64)):
    feat = t(rng.standard_normal((N, Cc, H, W)).astype(np.float32))
    boxes = np.concatenate([rng.uniform(-80, W * stride + 80, (N * H * W, 1)), rng.uniform(-80, H * stride + 80, (N * H * W, 1)),
                            rng.uniform(1, 12 * stride, (N * H * W, 2)), rng.uniform(-1.6, 0, (N * H * W, 1))], 1).astype(np.float32)
    for P in (1, 5):
        frm_forward(feat, t(boxes), 1.0 / stride, P); frm_backward(feat, t(boxes), 1.0 / stride, P)
torch.cuda.synchronize()
print('sanitize smoke ok')
