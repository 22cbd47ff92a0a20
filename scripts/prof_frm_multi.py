"""Development probe (GPU box): one multi-level FRM forward + backward per P, for ncu launch lists."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r3det_b200.fr import frm_forward_multi, frm_backward_multi
dev = torch.device('cuda:0')
rng = np.random.default_rng(4)
xs, bts, scales = [], [], []
for H, stride in ((128, 8), (64, 16), (32, 32), (16, 64), (8, 128)):
    xs.append(torch.randn((8, 256, H, H), device=dev))
    ys_, xs_ = np.meshgrid(np.arange(H) * stride, np.arange(H) * stride, indexing="ij")
    ctr = np.stack([xs_, ys_], -1).reshape(-1, 2).astype(np.float32)
    bx = np.zeros((8, H * H, 5), np.float32)
    bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (8, H * H, 2))
    bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (8, H * H, 2)))
    bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (8, H * H))
    bts.append(torch.from_numpy(bx.reshape(-1, 5)).to(dev)); scales.append(1.0 / stride)
for P in (1, 5):
    for _ in range(2):
        frm_forward_multi(xs, bts, scales, P); frm_backward_multi(xs, bts, scales, P)
torch.cuda.synchronize(); print('ok')
