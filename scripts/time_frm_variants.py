"""Development probe (GPU box): time gpurun_variants/frm_*.so (backward gather variants) on the batch-8, 5-level workload."""
import ctypes as C, glob, os, sys
import numpy as np, torch
dev = torch.device('cuda:0')
rng = np.random.default_rng(4)
xs, bts, scales, hw = [], [], [], []
for H, stride in ((128, 8), (64, 16), (32, 32), (16, 64), (8, 128)):
    xs.append(torch.randn((8, 256, H, H), device=dev))
    ys_, xs_ = np.meshgrid(np.arange(H) * stride, np.arange(H) * stride, indexing="ij")
    ctr = np.stack([xs_, ys_], -1).reshape(-1, 2).astype(np.float32)
    bx = np.zeros((8, H * H, 5), np.float32)
    bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (8, H * H, 2))
    bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (8, H * H, 2)))
    bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (8, H * H))
    bts.append(torch.from_numpy(bx.reshape(-1, 5)).to(dev)); scales.append(1.0 / stride); hw += [H, H]
outs = [torch.empty_like(x) for x in xs]
pa = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
hwc = (C.c_int * 10)(*hw); sc = (C.c_float * 5)(*scales)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for path in sorted(glob.glob('gpurun_variants/frm_*.so')):
    lib = C.CDLL(path)
    res = []
    for P in (1, 5):
        n = C.c_size_t(0)
        lib.r3g_frm_backward_multi_workspace_bytes(5, 8, hwc, P, C.byref(n))
        ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
        f = lambda: lib.r3g_frm_backward_multi_f32(5, pa(xs), pa(bts), 8, 256, hwc, sc, P, pa(outs), C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), st)
        g = lambda: lib.r3g_frm_forward_multi_f32(5, pa(xs), pa(bts), None, 8, 256, hwc, sc, P, pa(outs), st)
        for fn in (f, g):
            for _ in range(3): assert fn() == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(20): fn()
            e1.record(); e1.synchronize()
            res.append(e0.elapsed_time(e1) / 20)
    print(os.path.basename(path), 'P1 bwd %.3f fwd %.3f | P5 bwd %.3f fwd %.3f ms' % tuple(res), float(outs[0].sum()))
