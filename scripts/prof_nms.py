"""ncu target: one NMS call (v1, 15 classes, clustered) at K given on the command line."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r3det_b200 as R
from r3det_b200._nms_core import nms_device
from tests.util import clustered
dev = torch.device('cuda:0')
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
b, s, l = clustered(K, 2, 'v1')
B, S, L = (torch.from_numpy(x).to(dev) for x in (b, s, l))
sc = torch.tensor(float(b.max() + 1), device=dev)
for _ in range(3):
    keep, num = nms_device(B, S, 0.1, 'v1', labels=L, class_offset=sc, order_index=True)
torch.cuda.synchronize()
print('kept', int(num))
