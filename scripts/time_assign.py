import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r3det_b200 as R
from tests.util import rand_obb
dev = torch.device('cuda:0')
gt = torch.from_numpy(rand_obb(1000, 1)).to(dev); an = torch.from_numpy(rand_obb(200000, 1000)).to(dev)
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize(); return e0.elapsed_time(e1) / iters
for lq, aa in ((True, True), (True, False), (False, True)):
    print('match_low_quality', lq, 'assign_all', aa, '%.3f ms' % timeit(lambda: R.max_iou_assign(gt, an, 0.5, 0.4, 0.0, lq, aa, 'v1')))
print('matrix %.3f ms' % timeit(lambda: R.pairwise_iou(gt, an, 'v1')))
