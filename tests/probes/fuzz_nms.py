"""Randomised parity sweep (GPU box): batched / per-class / single-class NMS at random sizes against the oracle.
`run()` is also called by tests/test_nms_gpu.py with a fixed seed."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from r3det_b200._nms_core import nms_device
from oracle import port
from tests.util import clustered, rand_obb


def run(seed=0, iters=60, dev=None, verbose=True, large=False):
    """Returns the number of images whose keep list differs from the oracle's.  large=True: few classes and thousands of
    candidates per class, so that segments need several rounds of chunks (> 2048 rows) and batches cross the work threshold
    above which the rounds kernel switches to quarter chunks."""
    dev = torch.device('cuda:0') if dev is None else dev
    rng = np.random.default_rng(seed)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    bad = 0
    t0 = time.time()
    for it in range(iters):
        v = ['v1', 'v2', 'v3'][it % 3]
        nimg = int(rng.choice([1, 3, 6])) if large else int(rng.choice([1, 1, 2, 5, 9]))
        ncls = int(rng.choice([1, 2, 4])) if large else int(rng.choice([1, 2, 15, 40]))
        sizes = [int(rng.choice([2100, 9000, 30000] if large else [0, 1, 63, 64, 65, 300, 1000, 2500])) for _ in range(nimg)]
        parts = [clustered(max(k, 1), int(rng.integers(1 << 30)), v, ncls=ncls) for k in sizes]
        if large and rng.random() < 0.25:                     # sparse small boxes: (almost) nothing is suppressed, the worst case of the rounds
            parts = [(rand_obb(max(k, 1), int(rng.integers(1 << 30)), v, 2, 6), p[1], p[2]) for p, k in zip(parts, sizes)]
        parts = [(b[:k], s[:k], l[:k]) for (b, s, l), k in zip(parts, sizes)]
        b = np.concatenate([p[0] for p in parts]); s = np.concatenate([p[1] for p in parts]); l = np.concatenate([p[2] for p in parts])
        if len(b) == 0:
            continue
        ties = rng.random() < 0.3
        if ties:
            s[rng.integers(0, len(s), max(1, len(s) // 10))] = s[0]            # exact score ties
        cuts = np.cumsum([0] + sizes)
        parts = [(pb, s[cuts[i]:cuts[i + 1]], pl) for i, (pb, _, pl) in enumerate(parts)]   # the oracle sees the tied scores too
        bid = np.concatenate([np.full(k, i, np.int64) for i, k in enumerate(sizes)])
        thr = float(rng.choice([0.1, 0.3, 0.5]))
        by_index = bool(rng.integers(2))
        use_off = v != 'v2' and bool(rng.integers(2))
        scales = np.array([float(p[0].max() + 1) if len(p[0]) else 1.0 for p in parts], np.float32)
        kw = dict(labels=t(l), order_index=by_index, drop_small=(v == 'v3'), sort_path=bool(rng.integers(2)))
        if nimg > 1:
            kw.update(batch_ids=t(bid), n_batches=nimg)
        if use_off:
            kw['class_offset'] = t(scales) if nimg > 1 else torch.tensor(scales[0], device=dev)
        keep, num = nms_device(t(b), t(s), thr, v, **kw)
        keep = keep[:int(num.sum())].cpu().numpy()
        num = np.atleast_1d(num.cpu().numpy())
        start, off0 = 0, 0
        for i, (pb, ps, pl) in enumerate(parts):
            k = keep[start:start + int(num[i])] - off0
            start += int(num[i])
            want = np.zeros((0,), np.int64)
            if len(pb):
                bo = pb.copy()
                if use_off:                                                     # offsets separate the classes, as in the wrappers
                    o = (pl.astype(np.float32) * scales[i]).astype(np.float32)
                    bo[:, 0] += o; bo[:, 1] += o
                    want = port.nms(bo, ps, thr, v, inclusive=False)
                else:
                    want = port.nms(pb, ps, thr, v, labels=pl.astype(np.float32), inclusive=False)
                if by_index:
                    want = np.sort(want)
            if not np.array_equal(k, want):
                bad += 1
                if verbose:
                    print('MISMATCH it', it, v, 'img', i, 'K', len(pb), 'ncls', ncls, 'thr', thr, 'by_index', by_index, 'offsets', use_off,
                          'sort_path', kw['sort_path'], 'ties', ties, len(k), len(want), sorted(set(k.tolist()) ^ set(want.tolist()))[:6])
            off0 += len(pb)
    if verbose:
        print('fuzz done: %d mismatching images, %.1f s' % (bad, time.time() - t0))
    return bad


if __name__ == '__main__':
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 60,
                      large=len(sys.argv) > 3 and sys.argv[3] == 'large') else 0)
