import ctypes as C, glob, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import rand_obb
dev = torch.device('cuda:0')
a = torch.from_numpy(rand_obb(1000, 1, 'v1')).to(dev); b = torch.from_numpy(rand_obb(200000, 1000, 'v1')).to(dev)
out = torch.empty((1000, 200000), device=dev)
vp, i64, i32, sz = C.c_void_p, C.c_int64, C.c_int, C.c_size_t
for path in sorted(glob.glob('gpurun_variants/iou_*.so')):
    lib = C.CDLL(path)
    n = sz(0); lib.r3g_iou_workspace_bytes.argtypes = [i64, i64, C.POINTER(sz)]; lib.r3g_iou_workspace_bytes(1000, 200000, C.byref(n))
    ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
    f = lib.r3g_iou_matrix_prepared_f32; f.argtypes = [vp, i64, i64, vp, i64, i64, i32, i32, i32, vp, vp, sz, vp]
    g = lib.r3g_iou_matrix_f32; g.argtypes = f.argtypes
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (a.data_ptr(), 1000, 5, b.data_ptr(), 200000, 5, 1, 0, 1, out.data_ptr(), ws.data_ptr(), ws.numel(), st)
    assert g(*args) == 0
    for _ in range(5): f(*args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(50): f(*args)
    e1.record(); e1.synchronize()
    print(os.path.basename(path), '%.4f ms' % (e0.elapsed_time(e1) / 50), float(out.sum()))
