"""Development probe (GPU box): IoU parity vs oracle + timing vs the reference CUDA kernels."""
import ctypes as C, os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import r3det_b200
from r3det_b200.rbbox_geo import pairwise_iou, aligned_iou
from oracle import port

AR = {'v1': (-np.pi/2, 0), 'v2': (-np.pi/4, 3*np.pi/4), 'v3': (-np.pi/2, np.pi/2)}
def rand_obb(n, seed, ar, lo=8, hi=512):
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, 1024, n); cy = rng.uniform(0, 1024, n)
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n)); h = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    a = rng.uniform(*ar, n)
    return np.stack([cx, cy, w, h, a], 1).astype(np.float32)

def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    ts = []
    for _ in range(iters):
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))

dev = torch.device('cuda:0')
print(torch.cuda.get_device_name(0))
res = {}
for v in ['v1', 'v2', 'v3']:
    A = rand_obb(1000, 1, AR[v]); B = rand_obb(200000, 0, AR[v])
    a, b = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    out, stats = pairwise_iou(a, b, v, return_stats=True)
    torch.cuda.synchronize()
    st = stats.cpu().numpy().tolist()
    sub = out[:, :3000].cpu().numpy()
    o = port.iou_matrix(A, B[:3000], v, wrapper_mask=False)
    d = np.abs(sub - o)
    print(v, 'parity vs oracle (1000x3000): max %.3g n>1e-5 %d n>1e-6 %d' % (d.max(), (d > 1e-5).sum(), (d > 1e-6).sum()), 'stats', st)
    # no-workspace-alloc timing of the C call path (includes prep kernels + memset)
    med, best = timeit(lambda: pairwise_iou(a, b, v))
    print(v, '200k x 1k: median %.3f ms best %.3f ms -> %.1f Gpairs/s ; store GB/s %.0f' % (med, best, 2e8/med/1e6, 0.8/ (med*1e-3)))
    res[v] = dict(ms=med, best=best, stats=st, maxdiff=float(d.max()))
    al = aligned_iou(a, b[:1000], v).cpu().numpy()
    oa = port.iou_aligned(A, B[:1000], v)
    print(v, 'aligned max diff %.3g' % np.abs(al - oa).max())
    del out
# dense overlap regime
for v in ['v1']:
    rng = np.random.default_rng(5)
    centers = rng.uniform(100, 900, (32, 2))
    def dense(n, seed):
        r = np.random.default_rng(seed); b = rand_obb(n, seed, AR[v], 64, 256)
        idx = r.integers(0, 32, n); b[:, :2] = centers[idx] + r.normal(0, 20, (n, 2)); return b.astype(np.float32)
    A = dense(1000, 2); B = dense(200000, 3)
    a, b = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    out, stats = pairwise_iou(a, b, v, return_stats=True); torch.cuda.synchronize()
    med, best = timeit(lambda: pairwise_iou(a, b, v), iters=10)
    sub = out[:, :2000].cpu().numpy(); o = port.iou_matrix(A, B[:2000], v, wrapper_mask=False); d = np.abs(sub - o)
    print('dense', v, 'median %.3f ms -> %.1f Gpairs/s' % (med, 2e8/med/1e6), 'stats', stats.cpu().numpy().tolist(), 'max diff %.3g n>1e-5 %d' % (d.max(), (d > 1e-5).sum()))

# reference CUDA kernels
refdir = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'oracle', '_ref')
try:
    l1 = C.CDLL(os.path.join(refdir, 'libref_cuda_v1iou.so'))
    l1.refcuda_v1_iou_matrix.restype = C.c_float
    l1.refcuda_v1_iou_matrix.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int]
    A = rand_obb(1000, 1, AR['v1']); B = rand_obb(200000, 0, AR['v1'])
    a, b = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    ro = torch.empty((1000, 200000), device=dev)
    ms = l1.refcuda_v1_iou_matrix(a.data_ptr(), 1000, b.data_ptr(), 200000, 0, ro.data_ptr(), 3)
    ms = l1.refcuda_v1_iou_matrix(a.data_ptr(), 1000, b.data_ptr(), 200000, 0, ro.data_ptr(), 5)
    mine = pairwise_iou(a, b, 'v1')
    d = (mine - ro).abs()
    print('REF CUDA v1 kernel: %.3f ms -> %.2f Gpairs/s ; mine vs ref-cuda: max %.3g n>1e-5 %d of %d' % (ms, 2e8/ms/1e6, d.max().item(), (d > 1e-5).sum().item(), d.numel()))
    l3 = C.CDLL(os.path.join(refdir, 'libref_cuda_v3iou.so'))
    l3.refcuda_v3_iou_matrix.restype = C.c_float
    l3.refcuda_v3_iou_matrix.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int]
    A = rand_obb(1000, 1, AR['v3']); B = rand_obb(200000, 0, AR['v3'])
    a, b = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    ms = l3.refcuda_v3_iou_matrix(a.data_ptr(), 1000, b.data_ptr(), 200000, 1, ro.data_ptr(), 3)
    ms = l3.refcuda_v3_iou_matrix(a.data_ptr(), 1000, b.data_ptr(), 200000, 1, ro.data_ptr(), 5)
    mine = pairwise_iou(a, b, 'v3')
    d = (mine - ro).abs()
    print('REF CUDA v3 kernel: %.3f ms -> %.2f Gpairs/s ; mine vs ref-cuda: max %.3g n>1e-5 %d' % (ms, 2e8/ms/1e6, d.max().item(), (d > 1e-5).sum().item()))
except Exception as e:
    print('ref cuda failed:', repr(e))
