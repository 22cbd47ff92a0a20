"""Randomised parity sweep (GPU box): FRM forward / backward at random shapes against the oracle — widths that are and are not
multiples of 4 (TMA kernels / gather kernels), planes smaller than a tile, odd channel counts, boxes jittered by a fraction of
a cell up to several tiles (taps outside the staged window), boxes far outside the image, both point counts, single- and
multi-level calls.  `run()` returns the number of failing cases."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from r3det_b200.fr import frm_backward, frm_backward_multi, frm_forward, frm_forward_multi
from oracle import port


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def _level(rng, N, Cc, H, W, stride, jitter):
    feat = rng.standard_normal((N, Cc, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H) * stride, np.arange(W) * stride, indexing="ij")
    ctr = np.stack([xs, ys], -1).reshape(-1, 2).astype(np.float32)
    boxes = np.zeros((N, H * W, 5), np.float32)
    boxes[:, :, :2] = ctr[None] + rng.normal(0, stride * jitter, (N, H * W, 2))
    boxes[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (N, H * W, 2)))
    boxes[:, :, 4] = rng.uniform(-np.pi / 2, 0, (N, H * W))
    far = rng.random((N, H * W)) < 0.02
    boxes[far, :2] += rng.choice([-1.0, 1.0], (int(far.sum()), 2)) * stride * max(H, W) * 2
    return feat, boxes.reshape(-1, 5)


def run(seed=0, iters=40, dev=None, verbose=True):
    dev = torch.device('cuda:0') if dev is None else dev
    rng = np.random.default_rng(seed)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    bad, t0 = 0, time.time()
    for it in range(iters):
        N, Cc = int(rng.integers(1, 4)), int(rng.choice([1, 3, 8, 31, 32, 33, 70]))
        P = int(rng.choice([1, 5]))
        jitter = float(rng.choice([0.3, 1.0, 4.0, 20.0]))
        nl = int(rng.choice([1, 1, 3]))
        hw = [(int(rng.choice([1, 4, 7, 12, 16, 33, 40, 64])), int(rng.choice([1, 4, 7, 12, 16, 36, 44, 64]))) for _ in range(nl)]
        strides = [int(rng.choice([8, 16, 64])) for _ in range(nl)]
        lv = [_level(rng, N, Cc, H, W, s, jitter) for (H, W), s in zip(hw, strides)]
        gouts = [rng.standard_normal(f.shape).astype(np.float32) for f, _ in lv]
        if nl == 1:
            f, b = lv[0]
            got_f = [frm_forward(t(f), t(b), 1.0 / strides[0], P).cpu().numpy()]
            got_b = [frm_backward(t(gouts[0]), t(b), 1.0 / strides[0], P).cpu().numpy()]
        else:
            got_f = [y.cpu().numpy() for y in frm_forward_multi([t(f) for f, _ in lv], [t(b) for _, b in lv], [1.0 / s for s in strides], P)]
            got_b = [y.cpu().numpy() for y in frm_backward_multi([t(g) for g in gouts], [t(b) for _, b in lv], [1.0 / s for s in strides], P)]
        for k, ((f, b), g) in enumerate(zip(lv, gouts)):
            ef = _rel(got_f[k], port.frm_forward(f, b, 1.0 / strides[k], P))
            eb = _rel(got_b[k], port.frm_backward(g, b, 1.0 / strides[k], P, acc64=True))
            if not (ef <= 1e-5 and eb <= 1e-5):
                bad += 1
                if verbose:
                    print('MISMATCH it', it, 'level', k, (N, Cc) + hw[k], 'stride', strides[k], 'P', P, 'jitter', jitter, 'fwd', ef, 'bwd', eb)
    if verbose:
        print('frm fuzz done: %d failing levels, %.1f s' % (bad, time.time() - t0))
    return bad


if __name__ == '__main__':
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 40) else 0)
