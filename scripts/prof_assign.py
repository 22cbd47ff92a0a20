"""ncu target: the fused assigner on the headline shapes (1000 GT x 200k anchors, v1)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r3det_b200 as R
from tests.util import rand_obb
dev = torch.device("cuda:0")
gt = torch.from_numpy(rand_obb(1000, 1, "v1")).to(dev); an = torch.from_numpy(rand_obb(200000, 1000, "v1")).to(dev)
for _ in range(3):
    o = R.max_iou_assign(gt, an, 0.5, 0.4, 0.0, True, True, "v1")
torch.cuda.synchronize()
print("ok")
