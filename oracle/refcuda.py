"""TEST / BENCH INFRASTRUCTURE ONLY — ctypes access to the reference's UNMODIFIED CUDA kernels in oracle/_ref/libref_cuda_*.so
(built by oracle/build_ref.py from the sources under /root/reference; nothing is copied).  They are the "kernel to beat":
bench.py's `reference_cuda` leg times them on the same B200 in the same run, and tests/golden/make_golden_gpu.py uses the same
entry points to produce the *_refcuda.npz fixtures.  The product package never imports this module.

Every timing function returns the mean milliseconds per call over `iters` calls, measured with CUDA events inside the shim
(oracle/refshim/refcuda_common.h), after one untimed warm-up call."""
import ctypes as C
import os

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_libs = {}
_vp, _i64, _f32, _i32 = C.c_void_p, C.c_int64, C.c_float, C.c_int


def available(name):
    return os.path.exists(os.path.join(_DIR, name))


def _lib(name):
    if name not in _libs:
        import torch  # noqa: F401  (the shims link against libtorch)
        _libs[name] = C.CDLL(os.path.join(_DIR, name))
    return _libs[name]


def v1_iou_ms(b1, b2, out, iters=3, iof=False):
    """rbbox_geo mat_iou_iof_kernel (r3det/ops/rbbox_geo/src/rbbox_geo_kernel.cu:230-268): b1 (m,5), b2 (n,5), out (m,n) CUDA f32."""
    f = _lib("libref_cuda_v1iou.so").refcuda_v1_iou_matrix
    f.restype, f.argtypes = _f32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32]
    f(b1.data_ptr(), b1.size(0), b2.data_ptr(), b2.size(0), int(iof), out.data_ptr(), 1)
    return float(f(b1.data_ptr(), b1.size(0), b2.data_ptr(), b2.size(0), int(iof), out.data_ptr(), iters))


def v3_iou_ms(b1, b2, out, iters=3):
    """box_iou_rotated_cuda_kernel (r3det/ops/box_iou_rotated/src/box_iou_rotated_cuda.cu:13-63)."""
    f = _lib("libref_cuda_v3iou.so").refcuda_v3_iou_matrix
    f.restype, f.argtypes = _f32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32]
    f(b1.data_ptr(), b1.size(0), b2.data_ptr(), b2.size(0), 1, out.data_ptr(), 1)
    return float(f(b1.data_ptr(), b1.size(0), b2.data_ptr(), b2.size(0), 1, out.data_ptr(), iters))


def v1_nms_ms(dets6, thr, iters=2):
    """nmsr_cuda (r3det/ops/rnms/src/rcuda/rnms_kernel.cu:270-335: mask kernel + D2H + host scan): dets (K,6) CUDA f32.
    Returns (ms, kept count)."""
    import numpy as np
    f = _lib("libref_cuda_v1nms.so").refcuda_v1_nms
    f.restype, f.argtypes = _f32, [_vp, _i64, _f32, _vp, _vp, _i32]
    K = dets6.size(0)
    keep = np.empty((K,), np.int64); n = np.zeros((1,), np.int64)
    f(dets6.data_ptr(), K, float(thr), keep.ctypes.data, n.ctypes.data, 1)
    ms = float(f(dets6.data_ptr(), K, float(thr), keep.ctypes.data, n.ctypes.data, iters))
    return ms, int(n[0])


def v3_nms_ms(dets5, scores, thr, iters=2):
    """nms_rotated_cuda (r3det/ops/nms_rotated/src/nms_rotated_cuda.cu:71-134).  Returns (ms, kept count)."""
    import numpy as np
    f = _lib("libref_cuda_v3nms.so").refcuda_v3_nms
    f.restype, f.argtypes = _f32, [_vp, _vp, _i64, _f32, _vp, _vp, _i32]
    K = dets5.size(0)
    keep = np.empty((K,), np.int64); n = np.zeros((1,), np.int64)
    f(dets5.data_ptr(), scores.data_ptr(), K, float(thr), keep.ctypes.data, n.ctypes.data, 1)
    ms = float(f(dets5.data_ptr(), scores.data_ptr(), K, float(thr), keep.ctypes.data, n.ctypes.data, iters))
    return ms, int(n[0])


def poly_nms_ms(dets9, thr, iters=2):
    """poly_nms_cuda (r3det/ops/nms_rotated/src/poly_nms_cuda.cu:196-262).  Returns (ms, kept count)."""
    import numpy as np
    f = _lib("libref_cuda_polynms.so").refcuda_poly_nms
    f.restype, f.argtypes = _f32, [_vp, _i64, _f32, _vp, _vp, _i32]
    K = dets9.size(0)
    keep = np.empty((K,), np.int64); n = np.zeros((1,), np.int64)
    f(dets9.data_ptr(), K, float(thr), keep.ctypes.data, n.ctypes.data, 1)
    ms = float(f(dets9.data_ptr(), K, float(thr), keep.ctypes.data, n.ctypes.data, iters))
    return ms, int(n[0])


def frm_ms(feat, boxes, scale, points, out, backward=False, iters=3):
    """feature_refine_forward / backward kernels (r3det/ops/fr/src/feature_refine_kernel.cu:112-230) incl. the zero fill the
    reference's Python does first (feature_refine_module.py:21,36).  feat (N,C,H,W), boxes (N*H*W,5), out like feat."""
    lib = _lib("libref_cuda_frm.so")
    f = lib.refcuda_frm_backward if backward else lib.refcuda_frm_forward
    f.restype, f.argtypes = _f32, [_vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _i32]
    N, Cc, H, W = feat.shape
    f(feat.data_ptr(), boxes.data_ptr(), N, Cc, H, W, float(scale), int(points), out.data_ptr(), 1)
    return float(f(feat.data_ptr(), boxes.data_ptr(), N, Cc, H, W, float(scale), int(points), out.data_ptr(), iters))
