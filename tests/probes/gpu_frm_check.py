"""Development probe (GPU box): FRM + transforms parity vs oracle/reference CUDA, timing."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import r3det_b200
from r3det_b200.fr import frm_forward, frm_backward
from oracle import port
from oracle import transforms_np as tnp
from tests.util import rand_obb
dev = torch.device('cuda:0')
rng = np.random.default_rng(3)
def mk(N, Cc, H, W, stride):
    feat = rng.standard_normal((N, Cc, H, W)).astype(np.float32); gout = rng.standard_normal((N, Cc, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H) * stride, np.arange(W) * stride, indexing='ij')
    ctr = np.stack([xs, ys], -1).reshape(-1, 2).astype(np.float32)
    boxes = np.zeros((N, H * W, 5), np.float32)
    boxes[:, :, :2] = ctr[None] + rng.normal(0, stride, (N, H * W, 2))
    boxes[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (N, H * W, 2)))
    boxes[:, :, 4] = rng.uniform(-np.pi / 2, 0, (N, H * W))
    return feat, gout, boxes.reshape(-1, 5)
for (N, Cc, H, W, stride) in [(2, 8, 16, 16, 8), (1, 5, 9, 13, 16), (2, 40, 32, 32, 8)]:
    feat, gout, boxes = mk(N, Cc, H, W, stride)
    tf, tg, tb = (torch.from_numpy(x).to(dev) for x in (feat, gout, boxes))
    for P in (1, 5):
        o = frm_forward(tf, tb, 1.0 / stride, P).cpu().numpy(); want = port.frm_forward(feat, boxes, 1.0 / stride, P)
        gi = frm_backward(tg, tb, 1.0 / stride, P).cpu().numpy(); wantb = port.frm_backward(gout, boxes, 1.0 / stride, P, acc64=True)
        rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
        print((N, Cc, H, W), 'P', P, 'fwd rel %.2e bwd rel %.2e' % (rel(o, want), rel(gi, wantb)))
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
lf = None
try:
    lf = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'oracle', '_ref', 'libref_cuda_frm.so'))
    for fn in (lf.refcuda_frm_forward, lf.refcuda_frm_backward):
        fn.restype = C.c_float; fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int]
except Exception as e:
    print('no ref cuda frm', e)
tot = {1: [0, 0, 0, 0], 5: [0, 0, 0, 0]}
for H, stride in [(128, 8), (64, 16), (32, 32), (16, 64), (8, 128)]:
    feat, gout, boxes = mk(8, 256, H, H, stride)
    tf, tg, tb = (torch.from_numpy(x).to(dev) for x in (feat, gout, boxes))
    for P in (1, 5):
        f = timeit(lambda: frm_forward(tf, tb, 1.0 / stride, P)); b = timeit(lambda: frm_backward(tg, tb, 1.0 / stride, P))
        el = tf.numel()
        msg = 'H %3d P %d fwd %.3f ms (%.0f GB/s) bwd %.3f ms (%.0f GB/s)' % (H, P, f, el * 8 / f / 1e6, b, el * 8 / b / 1e6)
        tot[P][0] += f; tot[P][1] += b
        if lf is not None:
            o = torch.empty_like(tf)
            rf = lf.refcuda_frm_forward(tf.data_ptr(), tb.data_ptr(), 8, 256, H, H, 1.0 / stride, P, o.data_ptr(), 5)
            mine = frm_forward(tf, tb, 1.0 / stride, P)
            dfw = ((mine - o).abs().max() / o.abs().max()).item()
            rb = lf.refcuda_frm_backward(tg.data_ptr(), tb.data_ptr(), 8, 256, H, H, 1.0 / stride, P, o.data_ptr(), 5)
            mineb = frm_backward(tg, tb, 1.0 / stride, P)
            dbw = ((mineb - o).abs().max() / o.abs().max()).item()
            msg += ' | REF fwd %.3f ms bwd %.3f ms ; rel diff fwd %.1e bwd %.1e' % (rf, rb, dfw, dbw)
            tot[P][2] += rf; tot[P][3] += rb
        print(msg)
for P in (1, 5):
    print('batch-8 all levels P=%d: fwd %.3f ms bwd %.3f ms | ref fwd %.3f ms bwd %.3f ms ; roofline 8B/el: %.3f ms' % (P, *tot[P], 44695552 * 8 / 6453.7e6))
# transforms
for v in ('v1', 'v2', 'v3'):
    obb = rand_obb(5000, 4, v, 4, 600); t = torch.from_numpy(obb).to(dev)
    poly = r3det_b200.obb2poly(t, v)
    d = [np.abs(poly.cpu().numpy() - tnp.obb2poly(obb, v)).max(), np.abs(r3det_b200.obb2hbb(t, v).cpu().numpy() - tnp.obb2hbb(obb, v)).max(),
         np.abs(r3det_b200.obb2xyxy(t, v).cpu().numpy() - tnp.obb2xyxy(obb, v)).max(),
         np.abs(r3det_b200.poly2obb(poly, v).cpu().numpy()[:, :4] - tnp.poly2obb(poly.cpu().numpy(), v)[:, :4]).max()]
    print(v, 'transforms max abs diffs (obb2poly, obb2hbb, obb2xyxy, poly2obb[:4])', ['%.2e' % x for x in d])
