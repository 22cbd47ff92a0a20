import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r3det_b200.fr import frm_forward, frm_backward
dev = torch.device('cuda:0')
H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
P = int(sys.argv[2]) if len(sys.argv) > 2 else 1
stride = 1024 // H
rng = np.random.default_rng(4)
x = torch.randn((8, 256, H, H), device=dev)
ys, xs = np.meshgrid(np.arange(H) * stride, np.arange(H) * stride, indexing="ij")
ctr = np.stack([xs, ys], -1).reshape(-1, 2).astype(np.float32)
bx = np.zeros((8, H * H, 5), np.float32)
bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (8, H * H, 2))
bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (8, H * H, 2)))
bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (8, H * H))
bt = torch.from_numpy(bx.reshape(-1, 5)).to(dev)
for _ in range(3):
    y = frm_forward(x, bt, 1.0 / stride, P); g = frm_backward(x, bt, 1.0 / stride, P)
torch.cuda.synchronize(); print('ok')
