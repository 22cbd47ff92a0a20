"""GPU timing probe for the NMS path: K sweep (15 classes, clustered, v1 with class offsets), single class, batches of 8 images
and the worst case (nothing suppressed).  Prints one JSON line.  Tuning knobs are read by the library from the environment
(R3G_NMS_CHUNK, R3G_NMS_GRID), so run it once per setting."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r3det_b200 as R  # noqa: E402
from r3det_b200 import _lib as L  # noqa: E402
from r3det_b200._nms_core import nms_device  # noqa: E402
from tests.util import clustered, rand_obb  # noqa: E402

dev = torch.device("cuda:0")
LB = int(os.environ.get("LABEL_BITS", "4"))


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def graph_time(fn, iters=20):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timeit(g.replay, iters)


out = {"chunk": os.environ.get("R3G_NMS_CHUNK", "default"), "grid": os.environ.get("R3G_NMS_GRID", "default"), "sweep": {},
       "single_class": {}, "batch8": {}, "nothing_suppressed": {}, "workspace_mb": {}}
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "2000,8000,20000,80000,200000".split(","))]
for K in sizes:
    b, s, l = clustered(K, 2, "v1")
    B, S, Lb = (torch.from_numpy(x).to(dev) for x in (b, s, l))
    sc = torch.tensor(float(b.max() + 1), device=dev)
    fn = lambda: nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=sc, order_index=True, label_bits=LB)
    keep, num = fn()
    st = {}
    nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=sc, order_index=True, stats=st)
    torch.cuda.synchronize()
    c = st["counters"].cpu().numpy()
    ts = c[9:9 + int(c[8])]
    rec = {"ms": timeit(fn, 5 if K >= 80000 else 20), "kept": int(num), "pairs_stage1": int(c[0]), "pairs_sat": int(c[1]),
           "pairs_area": int(c[2]), "pairs_emu": int(c[3]), "rounds": int(c[4]), "items_mask": int(c[5]), "items_apply": int(c[6]),
           "phase_us": [round(float(x) / 1e3, 1) for x in np.diff(ts)]}
    try:
        rec["graph_ms"] = graph_time(fn)
    except Exception as e:  # noqa: BLE001
        rec["graph_ms"] = f"capture failed: {e}"
    out["sweep"][str(K)] = rec
    nb = C.c_size_t(0)
    L.check(L.lib().r3g_nms_workspace_bytes(K, C.byref(nb)))
    out["workspace_mb"][str(K)] = nb.value / 1e6
for K in (2000, 8000, 20000, 80000):
    b, s, _ = clustered(K, 2, "v1")
    B, S = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev)
    fn = lambda: nms_device(B, S, 0.1, "v1", order_index=True)
    keep, num = fn()
    out["single_class"][str(K)] = {"ms": timeit(fn, 10), "kept": int(num)}
for K in (2000, 8000, 20000, 80000, 200000):
    imgs = [clustered(K, 100 + i, "v1") for i in range(8)]
    B = torch.from_numpy(np.concatenate([x[0] for x in imgs])).to(dev)
    S = torch.from_numpy(np.concatenate([x[1] for x in imgs])).to(dev)
    Lb = torch.from_numpy(np.concatenate([x[2] for x in imgs])).to(dev)
    bid = torch.arange(8, device=dev).repeat_interleave(K)
    scales = torch.tensor([float(x[0].max() + 1) for x in imgs], device=dev)
    fn = lambda: nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=scales, order_index=True, batch_ids=bid, n_batches=8, label_bits=LB)
    keep, num = fn()
    st = {}
    nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=scales, order_index=True, batch_ids=bid, n_batches=8, stats=st, label_bits=LB)
    torch.cuda.synchronize()
    c = st["counters"].cpu().numpy()
    ms = timeit(fn, 5)
    out["batch8"][str(K)] = {"ms": ms, "kept": int(num.sum()), "mcands_per_s": 8 * K / ms / 1e3, "rounds": int(c[4]),
                             "pairs_stage1": int(c[0]), "phase_us": [round(float(x) / 1e3, 1) for x in np.diff(c[9:9 + int(c[8])])]}
    del B, S, Lb, bid
# worst case for the rounds: sparse small boxes, (almost) nothing suppressed -> the blocked triangular sweep
for K in (20000, 100000):
    b = rand_obb(K, 7, "v1", 2, 6)
    s = np.random.default_rng(8).permutation(np.linspace(0.05, 1, K)).astype(np.float32)
    B, S = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev)
    fn = lambda: nms_device(B, S, 0.1, "v1", order_index=True)
    keep, num = fn()
    out["nothing_suppressed"][str(K)] = {"ms": timeit(fn, 3), "kept": int(num)}
print(json.dumps(out))
