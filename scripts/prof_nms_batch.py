"""Development probe (GPU box): batched NMS (8 images x K candidates) a few times, for ncu launch lists."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r3det_b200._nms_core import nms_device
from tests.util import clustered
dev = torch.device('cuda:0')
K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
imgs = [clustered(K, 100 + i, 'v1') for i in range(8)]
B = torch.from_numpy(np.concatenate([x[0] for x in imgs])).to(dev); S = torch.from_numpy(np.concatenate([x[1] for x in imgs])).to(dev)
Lb = torch.from_numpy(np.concatenate([x[2] for x in imgs])).to(dev)
bid = torch.arange(8, device=dev).repeat_interleave(K); sc = torch.tensor([float(x[0].max() + 1) for x in imgs], device=dev)
for _ in range(3):
    keep, num = nms_device(B, S, 0.1, 'v1', labels=Lb, class_offset=sc, order_index=True, batch_ids=bid, n_batches=8)
torch.cuda.synchronize(); print('kept', int(num.sum()))
if len(sys.argv) > 2:                      # same candidates in shuffled order: the image-range shortcut cannot fire
    perm = torch.randperm(8 * K, device=dev)
    for _ in range(3):
        nms_device(B[perm], S[perm], 0.1, 'v1', labels=Lb[perm], class_offset=sc, order_index=True, batch_ids=bid[perm], n_batches=8)
    torch.cuda.synchronize()
