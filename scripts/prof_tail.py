"""Development probe (GPU box): the dense-head tail of bench.py (batch 8) — `get_bboxes_padded` twice, for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import r3det_b200 as R
dev = torch.device('cuda:0')
rng = np.random.default_rng(9)
Bn, A, Cn = 8, 9, 15
cls, reg, anc = [], [], []
for H, stride in ((128, 8), (64, 16), (32, 32), (16, 64), (8, 128)):
    c = rng.normal(-4.6, 1.0, (Bn, A * Cn, H, H)).astype(np.float32)
    hot = rng.random((Bn, A * Cn, H, H)) < 2e-4
    c[hot] = rng.normal(1.0, 1.0, int(hot.sum())).astype(np.float32)
    cls.append(torch.from_numpy(c).to(dev))
    reg.append(torch.from_numpy(rng.normal(0, 0.2, (Bn, A * 5, H, H)).astype(np.float32)).to(dev))
    ys, xs = np.meshgrid(np.arange(H), np.arange(H), indexing="ij")
    ctr = (np.stack([xs, ys], -1).reshape(-1, 1, 2) * stride + stride / 2).astype(np.float32)
    wh = (stride * 4 * np.array([[1, 1], [1.4, 0.7], [0.7, 1.4]], np.float32)[None].repeat(3, 1)
          * np.array([1, 1, 1, 1.26, 1.26, 1.26, 1.59, 1.59, 1.59], np.float32)[None, :, None])
    a = np.concatenate([np.broadcast_to(ctr, (H * H, A, 2)), np.broadcast_to(wh, (H * H, A, 2)), np.zeros((H * H, A, 1), np.float32)], -1)
    anc.append(torch.from_numpy(np.ascontiguousarray(a.reshape(-1, 5))).to(dev))
coder = R.DeltaXYWHAOBBoxCoder((0.,) * 5, (1.,) * 5, angle_range="v1")
metas = [dict(img_shape=(1024, 1024, 3), scale_factor=np.ones(4, np.float32))] * Bn
cfg = dict(nms_pre=2000, min_bbox_size=0, score_thr=0.05, nms=dict(type="v1", iou_thr=0.1), max_per_img=2000)
mode = sys.argv[1] if len(sys.argv) > 1 else "padded"
fn = (lambda: R.get_bboxes_padded(cls, reg, anc, metas, cfg, coder)) if mode == "padded" else (lambda: R.get_bboxes(cls, reg, anc, metas, cfg, coder))
fn(); torch.cuda.synchronize()
torch.cuda.nvtx.range_push("measured")
out = fn()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print(mode, "detections", int(out[2].sum()) if mode == "padded" else sum(d[0].size(0) for d in out))
