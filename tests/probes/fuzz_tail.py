"""Randomised sweep (GPU box): the synchronisation-free padded multiclass NMS of a batch against the per-image reference-API
call — random batch sizes (incl. > 64: split), rows, class counts, score thresholds (incl. images without any candidate),
max_num, every nms.type.  `run()` returns the number of differing images."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import r3det_b200 as R
from tests.util import clustered


def run(seed=0, iters=60, dev=None, verbose=True):
    dev = torch.device('cuda:0') if dev is None else dev
    rng = np.random.default_rng(seed)
    bad, t0 = 0, time.time()
    for it in range(iters):
        kind = ['v1', 'v3', 'v2', 'mmcv'][it % 4]
        geom = {'v1': 'v1', 'v3': 'v3'}.get(kind, 'v2')
        B = int(rng.choice([1, 2, 5, 8, 70]))
        n = int(rng.choice([1, 37, 300, 1200]))
        nc = int(rng.choice([1, 3, 15, 16, 17]))
        max_num = int(rng.choice([1, 20, 2000]))
        thr = float(rng.choice([0.05, 0.3, 0.9]))
        boxes = np.stack([clustered(n, int(rng.integers(1 << 30)), geom)[0] for _ in range(B)])
        scores = (rng.random((B, n, nc + 1)) ** int(rng.choice([1, 4, 12]))).astype(np.float32)
        if B > 1:
            scores[int(rng.integers(B))] = 0.0                          # an image without candidates
        mb = torch.from_numpy(boxes).to(dev); ms = torch.from_numpy(scores).to(dev)
        cfg = dict(type=kind, iou_thr=float(rng.choice([0.1, 0.5])))
        dets, labels, counts = R.multiclass_nms_rotated_padded(mb, ms, thr, cfg, max_num)
        cl = counts.cpu().numpy()
        for b in range(B):
            out = R.multiclass_nms_rotated(mb[b], ms[b], thr, cfg, max_num)
            d, l = out[0], out[1]
            c = int(cl[b])
            ok = c == d.size(0) and torch.equal(dets[b, :c], d) and torch.equal(labels[b, :c], l) and float(dets[b, c:].abs().sum()) == 0.0
            if not ok:
                bad += 1
                if verbose:
                    print('MISMATCH it', it, kind, 'B', B, 'n', n, 'nc', nc, 'max', max_num, 'thr', thr, 'img', b, c, d.size(0))
    if verbose:
        print('tail fuzz done: %d differing images, %.1f s' % (bad, time.time() - t0))
    return bad


if __name__ == '__main__':
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 60) else 0)
