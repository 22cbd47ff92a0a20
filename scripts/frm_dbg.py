import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import r3det_b200 as R
from r3det_b200.fr import frm_forward
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
shapes = {'a': (1, 4, 8, 8, 64), 'b': (2, 32, 32, 32, 32), 'c': (1, 8, 128, 128, 8)}
for (N, C, H, W, stride) in [shapes[sys.argv[1]]]:
    x = torch.randn((N, C, H, W), device=dev)
    ys_, xs_ = np.meshgrid(np.arange(H) * stride, np.arange(W) * stride, indexing="ij")
    ctr = np.stack([xs_, ys_], -1).reshape(-1, 2).astype(np.float32)
    bx = np.zeros((N, H * W, 5), np.float32)
    bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (N, H * W, 2))
    bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (N, H * W, 2)))
    bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (N, H * W))
    b = torch.from_numpy(bx.reshape(-1, 5)).to(dev)
    for P in (1, 5):
        os.environ["R3G_FRM_TMA"] = "1"
        y = frm_forward(x, b, 1.0 / stride, P)
        torch.cuda.synchronize()
        print("shape", (N, C, H, W), "P", P, "ok", float(y.sum()))
