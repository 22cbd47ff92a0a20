"""GPU timing probe for the config-3 IoU step: the pair kernel alone (prepared boxes) and the whole op, v1 / v3 / IoF, plus the
fused assigner.  Prints one JSON line."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r3det_b200 as R
from r3det_b200 import _lib as L
from tests.util import rand_obb
dev = torch.device("cuda:0")
lib = L.lib()
GT, AN = 1000, 200000
out = torch.empty((GT, AN), device=dev)


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / iters


res = {}
for v, mode in (("v1", 0), ("v3", 0), ("v1", 1)):
    gt = torch.from_numpy(rand_obb(GT, 1, v)).to(dev); an = torch.from_numpy(rand_obb(AN, 1000, v)).to(dev)
    nb = C.c_size_t(0)
    L.check(lib.r3g_iou_workspace_bytes(GT, AN, C.byref(nb)))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    full = lambda: L.check(lib.r3g_iou_matrix_f32(L.ptr(gt), GT, 5, L.ptr(an), AN, 5, L.V[v], mode, L.FLAG_STRICT, L.ptr(out), L.ptr(ws), ws.numel(), sp))
    kern = lambda: L.check(lib.r3g_iou_matrix_prepared_f32(L.ptr(gt), GT, 5, L.ptr(an), AN, 5, L.V[v], mode, L.FLAG_STRICT, L.ptr(out), L.ptr(ws), ws.numel(), sp))
    full()
    res[f"{v}_{'iof' if mode else 'iou'}"] = {"op_ms": timeit(full), "pair_kernel_ms": timeit(kern), "checksum": float(out.double().sum())}
gt = torch.from_numpy(rand_obb(GT, 1, "v1")).to(dev); an = torch.from_numpy(rand_obb(AN, 1000, "v1")).to(dev)
res["assign_ms"] = timeit(lambda: R.max_iou_assign(gt, an, 0.5, 0.4, 0.0, True, True, "v1"))
res["frac_of_hbm_6453.7"] = 0.8 / res["v1_iou"]["pair_kernel_ms"] / 6.4537
print(json.dumps(res))
