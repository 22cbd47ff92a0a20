import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import r3det_b200 as R
from r3det_b200 import _lib as L
from oracle import port
from tests.util import rand_obb, clustered
dev = torch.device('cuda:0')
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
for v in ['v1', 'v2', 'v3']:
    a, b = rand_obb(1000, 11, v), rand_obb(3000, 12, v)
    got = R.pairwise_iou(t(a), t(b), v, flags=L.FLAG_EMULATE_ALL).cpu().numpy()
    want = port.iou_matrix(a, b, v, wrapper_mask=False)
    d = np.abs(got - want); print(v, 'emulate-all vs oracle random: max %.3g n>1e-6 %d' % (d.max(), (d > 1e-6).sum()))
    c, _, _ = clustered(1500, 17, v)
    d0 = c[:600].copy(); e = d0.copy(); e[:, :2] += np.random.default_rng(0).normal(0, 1e-3, (600, 2)).astype(np.float32)
    for flags, name in [(L.FLAG_STRICT, 'strict'), (L.FLAG_EMULATE_ALL, 'emu-all'), (0, 'geometry')]:
        got = R.pairwise_iou(t(d0), t(e), v, flags=flags).cpu().numpy()
        want = port.iou_matrix(d0, e, v, wrapper_mask=False); ex = port.iou_exact(d0, e, v)
        d = np.abs(got - want)
        print(v, 'near-identical', name, 'vs oracle: max %.3g n>1e-5 %d of %d overlapping | oracle vs exact n>1e-5 %d | got vs exact n>1e-5 %d' % (
            d.max(), (d > 1e-5).sum(), (ex > 0).sum(), (np.abs(want - ex) > 1e-5).sum(), (np.abs(got - ex) > 1e-5).sum()))
