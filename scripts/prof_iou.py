"""ncu target: a few launches of the config-3 IoU step (v1, 1000 x 200000)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r3det_b200 as R
from tests.util import rand_obb
dev = torch.device('cuda:0')
v = sys.argv[1] if len(sys.argv) > 1 else 'v1'
a = torch.from_numpy(rand_obb(1000, 1, v)).to(dev); b = torch.from_numpy(rand_obb(200000, 1000, v)).to(dev)
for _ in range(4):
    o = R.pairwise_iou(a, b, v)
torch.cuda.synchronize()
print('done', float(o.max()))
