"""compute-sanitizer target: small invocations of every kernel (odd sizes, tails, empty inputs)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
torch.cuda.synchronize()
print('sanitize smoke ok')
