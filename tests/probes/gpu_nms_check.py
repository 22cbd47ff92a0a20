"""Development probe (GPU box): NMS parity vs oracle + timing."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import r3det_b200
from r3det_b200._nms_core import nms_device
from oracle import port

AR = {'v1': (-np.pi/2, 0), 'v2': (-np.pi/4, 3*np.pi/4), 'v3': (-np.pi/2, np.pi/2)}
def rand_obb(n, rng, ar, lo=8, hi=512):
    cx = rng.uniform(0, 1024, n); cy = rng.uniform(0, 1024, n)
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n)); h = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    a = rng.uniform(*ar, n)
    return np.stack([cx, cy, w, h, a], 1).astype(np.float32)
def clustered(K, seed, ar, ncls=15):
    rng = np.random.default_rng(seed)
    seeds = rand_obb(max(K // 10, 1), rng, ar, 12, 200)
    idx = rng.integers(0, len(seeds), K)
    b = seeds[idx].copy()
    b[:, 0:2] += rng.normal(0, 4, (K, 2)); b[:, 4] += rng.normal(0, 0.05, K); b[:, 2:4] *= np.exp(rng.normal(0, 0.1, (K, 2)))
    labels = (idx % ncls).astype(np.int64)
    scores = rng.permutation(np.linspace(0.05, 1, K)).astype(np.float32)
    return b.astype(np.float32), scores, labels

dev = torch.device('cuda:0')
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

for v in ['v1', 'v2', 'v3']:
    for K in [300, 2000, 8000]:
        b, s, l = clustered(K, 3, AR[v])
        bt, st_, lt = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(l).to(dev)
        # single class, score order
        keep, num = nms_device(bt, st_, 0.1, v); k = keep[:int(num)].cpu().numpy()
        ko = port.nms(b, s, 0.1, v, inclusive=False)
        ok1 = np.array_equal(k, ko)
        # inclusive + index order
        keep, num = nms_device(bt, st_, 0.1, v, inclusive=True, order_index=True); k2 = keep[:int(num)].cpu().numpy()
        ko2 = np.sort(port.nms(b, s, 0.1, v, inclusive=True))
        ok2 = np.array_equal(k2, ko2)
        # multi-label, no offsets (ml semantics)
        keep, num = nms_device(bt, st_, 0.1, v, labels=lt); k3 = keep[:int(num)].cpu().numpy()
        ko3 = port.nms(b, s, 0.1, v, labels=l.astype(np.float32), inclusive=False)
        ok3 = np.array_equal(k3, ko3)
        # with class offsets (batched wrappers)
        scale = np.float32(b.max() + 1)
        off = (l.astype(np.float32) * scale).astype(np.float32)
        bo = b.copy(); bo[:, 0] = bo[:, 0] + off; bo[:, 1] = bo[:, 1] + off
        keep, num = nms_device(bt, st_, 0.1, v, labels=lt, class_offset=torch.tensor(scale, device=dev), order_index=(v == 'v1'))
        k4 = keep[:int(num)].cpu().numpy()
        ko4 = port.nms(bo, s, 0.1, v, inclusive=False)
        if v == 'v1': ko4 = np.sort(ko4)
        ok4 = np.array_equal(k4, ko4)
        print(v, 'K', K, 'single', ok1, len(k), len(ko), '| incl+idx', ok2, '| labels', ok3, len(k3), len(ko3), '| offsets', ok4, len(k4), len(ko4))
        if not ok4:
            print('   symdiff', sorted(set(k4.tolist()) ^ set(ko4.tolist()))[:10])
        if not ok1:
            print('   symdiff single', sorted(set(k.tolist()) ^ set(ko.tolist()))[:10])

# wrappers
b, s, l = clustered(3000, 5, AR['v1'])
bt, st_, lt = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(l).to(dev)
d1, k1 = r3det_b200.batched_rnms(bt, st_, lt, 0.1)
d3, k3 = r3det_b200.obb_batched_nms(bt, st_, lt, 0.1)
k2 = r3det_b200.ml_nms_rotated(bt, st_, lt, 0.1)
dr, kr = r3det_b200.rnms(torch.cat([bt, st_[:, None]], 1), 0.1)
dn, kn = r3det_b200.rnms(np.concatenate([b, s[:, None]], 1), 0.1)
print('wrappers:', d1.shape, k1[:5].tolist(), d3.shape, k3[:5].tolist(), k2.shape, dr.shape, type(dn), dn.shape, 'numpy(>=) vs oracle', np.array_equal(kn, np.sort(port.nms(b, s, 0.1, 'v1', inclusive=True))))

# timing sweep (15 classes, clustered)
for K in [2000, 8000, 20000, 80000, 200000]:
    b, s, l = clustered(K, 2, AR['v1'])
    bt, st_, lt = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(l).to(dev)
    sc = torch.tensor(float(b.max() + 1), device=dev)
    for v in ['v1', 'v3']:
        f = lambda: nms_device(bt, st_, 0.1, v, labels=lt, class_offset=sc, order_index=(v == 'v1'))
        keep, num = f(); torch.cuda.synchronize()
        ms = timeit(f, iters=5, warm=2)
        print('K %6d %s 15cls: %.3f ms -> %.2f Mcands/s  kept %d' % (K, v, ms, K / ms / 1e3, int(num)))
for K in [2000, 8000, 20000]:
    b, s, l = clustered(K, 2, AR['v1'])
    bt, st_ = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev)
    f = lambda: nms_device(bt, st_, 0.1, 'v1', order_index=True)
    keep, num = f(); torch.cuda.synchronize()
    ms = timeit(f, iters=5, warm=2)
    print('K %6d v1 single-class: %.3f ms -> %.2f Mcands/s kept %d' % (K, ms, K / ms / 1e3, int(num)))
