"""Development probe (GPU box): polygon NMS + v1 NMS vs the reference's own CUDA kernels (oracle/_ref), parity and timing."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import r3det_b200 as R
from r3det_b200.nms_rotated import poly_nms_device
from oracle import port, transforms_np as T
from tests.util import clustered
dev = torch.device('cuda:0')
REFDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'oracle', '_ref')
vp, i64, f32, i32 = C.c_void_p, C.c_int64, C.c_float, C.c_int
lp = C.CDLL(os.path.join(REFDIR, 'libref_cuda_polynms.so'))
lp.refcuda_poly_nms.restype = f32; lp.refcuda_poly_nms.argtypes = [vp, i64, f32, vp, vp, i32]
lp.refcuda_poly_iou.restype = None; lp.refcuda_poly_iou.argtypes = [vp, vp, i64, vp]
lv = C.CDLL(os.path.join(REFDIR, 'libref_cuda_v1nms.so'))
lv.refcuda_v1_nms.restype = f32; lv.refcuda_v1_nms.argtypes = [vp, i64, f32, vp, vp, i32]

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize(); return e0.elapsed_time(e1) / iters

for K, jit, thr in ((1500, 0.0, 0.1), (900, 3.0, 0.3), (8000, 0.0, 0.1), (20000, 2.0, 0.1)):
    b, s, _ = clustered(K, 3, 'v1')
    p = T.obb2poly(b, 'v1')
    if jit: p = p + np.random.default_rng(1).normal(0, jit, p.shape).astype(np.float32)
    d = np.ascontiguousarray(np.concatenate([p, s[:, None]], 1), np.float32)
    t = torch.from_numpy(d).to(dev)
    keep = np.empty((K,), np.int64); n = np.zeros((1,), np.int64)
    ms_ref = lp.refcuda_poly_nms(t.data_ptr(), K, thr, keep.ctypes.data, n.ctypes.data, 3)
    kref = keep[:int(n[0])]
    _, k = R.poly_nms(t, thr)
    ms = timeit(lambda: poly_nms_device(t[:, :8], t[:, 8], thr))
    ko = port.poly_nms(d, thr) if K <= 2000 else None
    # device IoU of neighbouring rows: ours (via a 2-box NMS is awkward) -> compare oracle vs reference values instead
    o = torch.empty((K - 1,), device=dev)
    pa, qa = t[:-1, :8].contiguous(), t[1:, :8].contiguous()
    lp.refcuda_poly_iou(pa.data_ptr(), qa.data_ptr(), K - 1, o.data_ptr())
    po = port.poly_iou_aligned(d[:-1], d[1:])
    print('poly K %5d jitter %.0f: ours==ref %s (%d vs %d kept) oracle==ref %s | ours %.3f ms  ref CUDA %.3f ms | oracle-vs-ref IoU max diff %.2e' % (
        K, jit, np.array_equal(k.cpu().numpy(), kref), len(k), len(kref), None if ko is None else np.array_equal(ko, kref), ms, ms_ref,
        float(np.abs(po - o.cpu().numpy()).max())))
for K in (1500, 8000, 20000):
    b, s, _ = clustered(K, 6, 'v1')
    d = np.ascontiguousarray(np.concatenate([b, s[:, None]], 1), np.float32)
    t = torch.from_numpy(d).to(dev)
    keep = np.empty((K,), np.int64); n = np.zeros((1,), np.int64)
    ms_ref = lv.refcuda_v1_nms(t.data_ptr(), K, 0.1, keep.ctypes.data, n.ctypes.data, 3)
    kref = keep[:int(n[0])]
    _, k = R.rnms(t, 0.1)
    ms = timeit(lambda: R.rnms(t, 0.1))
    print('v1 rnms K %5d: ours==ref CUDA %s (%d vs %d) | ours %.3f ms (incl. wrapper sync)  ref CUDA %.3f ms' % (
        K, np.array_equal(k.cpu().numpy(), kref), len(k), len(kref), ms, ms_ref))
