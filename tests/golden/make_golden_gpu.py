"""Generate tests/golden/*_refcuda.npz from the reference's own CUDA kernels (run on a GPU box).

The v1 IoU op (rbbox_geo) and FRM (fr) have NO CPU implementation in the reference; their unmodified .cu files
are compiled for sm_100 into oracle/_ref/libref_cuda_*.so (oracle/build_ref.py --cuda) and driven here.

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out'     # then copy the .npz into tests/golden/
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.util import clustered, rand_obb  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")
out_dir = sys.argv[1] if len(sys.argv) > 1 else HERE
dev = torch.device("cuda:0")
vp, i64, f32, i32 = C.c_void_p, C.c_int64, C.c_float, C.c_int


def main():
    l1 = C.CDLL(os.path.join(REFDIR, "libref_cuda_v1iou.so"))
    l1.refcuda_v1_iou_matrix.restype = f32
    l1.refcuda_v1_iou_matrix.argtypes = [vp, i64, vp, i64, i32, vp, i32]
    l1.refcuda_v1_iou_aligned.restype = f32
    l1.refcuda_v1_iou_aligned.argtypes = [vp, i64, vp, i64, i32, vp, i32]
    g = {}
    c, _, _ = clustered(200, 5, "v1")
    b1 = np.concatenate([rand_obb(150, 1, "v1"), rand_obb(50, 2, "v1", 0.5, 8, 64), c[:100]]).astype(np.float32)
    b2 = np.concatenate([rand_obb(250, 3, "v1"), rand_obb(50, 4, "v1", 0.5, 8, 64), c[100:] + np.float32(0.5)]).astype(np.float32)
    t1, t2 = torch.from_numpy(b1).to(dev), torch.from_numpy(b2).to(dev)
    for iof in (0, 1):
        o = torch.empty((len(b1), len(b2)), device=dev)
        l1.refcuda_v1_iou_matrix(t1.data_ptr(), len(b1), t2.data_ptr(), len(b2), iof, o.data_ptr(), 1)
        g["v1_iof" if iof else "v1_iou"] = o.cpu().numpy()
    o = torch.empty((300,), device=dev)
    l1.refcuda_v1_iou_aligned(t1.data_ptr(), 300, t2[:300].contiguous().data_ptr(), 300, 0, o.data_ptr(), 1)
    g["v1_aligned"] = o.cpu().numpy()
    g["b1"], g["b2"] = b1, b2
    np.savez_compressed(os.path.join(out_dir, "iou_refcuda.npz"), **g)

    lf = C.CDLL(os.path.join(REFDIR, "libref_cuda_frm.so"))
    for fn in (lf.refcuda_frm_forward, lf.refcuda_frm_backward):
        fn.restype = f32
        fn.argtypes = [vp, vp, i32, i32, i32, i32, f32, i32, vp, i32]
    f = {}
    rng = np.random.default_rng(9)
    for tag, (N, Cc, H, W, stride) in {"a": (2, 6, 12, 10, 8), "b": (1, 4, 8, 8, 64)}.items():
        feat = rng.standard_normal((N, Cc, H, W)).astype(np.float32)
        gout = rng.standard_normal((N, Cc, H, W)).astype(np.float32)
        ys, xs = np.meshgrid(np.arange(H) * stride, np.arange(W) * stride, indexing="ij")
        ctr = np.stack([xs, ys], -1).reshape(-1, 2).astype(np.float32)
        boxes = np.zeros((N, H * W, 5), np.float32)
        boxes[:, :, :2] = ctr[None] + rng.normal(0, stride, (N, H * W, 2))
        boxes[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (N, H * W, 2)))
        boxes[:, :, 4] = rng.uniform(-np.pi / 2, 0, (N, H * W))
        boxes[0, :5, :2] = [[-50, -50], [1e4, 3], [3, 1e4], [-1.0 * stride, 2.0], [0, 0]]      # out-of-range / edge samples
        boxes = boxes.reshape(-1, 5)
        tf, tg, tb = (torch.from_numpy(x).to(dev) for x in (feat, gout, boxes))
        for P in (1, 5):
            o = torch.empty_like(tf); gi = torch.empty_like(tf)
            lf.refcuda_frm_forward(tf.data_ptr(), tb.data_ptr(), N, Cc, H, W, 1.0 / stride, P, o.data_ptr(), 1)
            lf.refcuda_frm_backward(tg.data_ptr(), tb.data_ptr(), N, Cc, H, W, 1.0 / stride, P, gi.data_ptr(), 1)
            f[f"{tag}_fwd_p{P}"] = o.cpu().numpy(); f[f"{tag}_bwd_p{P}"] = gi.cpu().numpy()
        f[f"{tag}_feat"], f[f"{tag}_gout"], f[f"{tag}_boxes"] = feat, gout, boxes
        f[f"{tag}_scale"] = np.float32(1.0 / stride)
    np.savez_compressed(os.path.join(out_dir, "frm_refcuda.npz"), **f)
    poly_and_v1nms()
    print("wrote", os.listdir(out_dir))


def quads(K, seed, jitter=0.0):
    """polygons of clustered v1 boxes (obb2poly_v1 corner order), optionally perturbed into general quadrilaterals"""
    from oracle import transforms_np as T
    b, s, _ = clustered(K, seed, "v1")
    p = T.obb2poly(b, "v1")
    if jitter:
        p = p + np.random.default_rng(seed).normal(0, jitter, p.shape).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([p, s[:, None]], 1), np.float32)


def poly_and_v1nms():
    """poly_nms_cuda.cu and rnms_kernel.cu, compiled unmodified (THC stand-in: oracle/refshim/thc)."""
    lp = C.CDLL(os.path.join(REFDIR, "libref_cuda_polynms.so"))
    lp.refcuda_poly_nms.restype = f32
    lp.refcuda_poly_nms.argtypes = [vp, i64, f32, vp, vp, i32]
    lp.refcuda_poly_iou.restype = None
    lp.refcuda_poly_iou.argtypes = [vp, vp, i64, vp]
    g = {}
    for tag, (K, seed, jit, thr) in {"rect": (1500, 3, 0.0, 0.1), "quad": (900, 4, 3.0, 0.3)}.items():
        d = quads(K, seed, jit)
        t = torch.from_numpy(d).to(dev)
        keep = np.empty((K,), np.int64); n = np.zeros((1,), np.int64)
        lp.refcuda_poly_nms(t.data_ptr(), K, thr, keep.ctypes.data, n.ctypes.data, 1)
        g[f"{tag}_dets"], g[f"{tag}_thr"], g[f"{tag}_keep"] = d, np.float32(thr), keep[:int(n[0])].copy()
        # pairwise values for spatial neighbours (sorted by the first vertex: most of these pairs overlap)
        order = np.lexsort((d[:, 1], np.round(d[:, 0] / 8)))
        pairs = np.stack([order[:-1], order[1:]], 1)
        p_, q_ = t[torch.from_numpy(pairs[:, 0]).to(dev), :8].contiguous(), t[torch.from_numpy(pairs[:, 1]).to(dev), :8].contiguous()
        o = torch.empty((K - 1,), device=dev)
        lp.refcuda_poly_iou(p_.data_ptr(), q_.data_ptr(), K - 1, o.data_ptr())
        g[f"{tag}_pairs"], g[f"{tag}_iou_pairs"] = pairs, o.cpu().numpy()
    lv = C.CDLL(os.path.join(REFDIR, "libref_cuda_v1nms.so"))
    lv.refcuda_v1_nms.restype = f32
    lv.refcuda_v1_nms.argtypes = [vp, i64, f32, vp, vp, i32]
    b, s_, _ = clustered(1500, 6, "v1")
    d = np.ascontiguousarray(np.concatenate([b, s_[:, None]], 1), np.float32)
    t = torch.from_numpy(d).to(dev)
    keep = np.empty((1500,), np.int64); n = np.zeros((1,), np.int64)
    lv.refcuda_v1_nms(t.data_ptr(), 1500, 0.1, keep.ctypes.data, n.ctypes.data, 1)
    g["v1nms_dets"], g["v1nms_keep"] = d, keep[:int(n[0])].copy()
    np.savez_compressed(os.path.join(out_dir, "poly_refcuda.npz"), **g)


if __name__ == "__main__":
    main()
