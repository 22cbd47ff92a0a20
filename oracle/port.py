"""TEST INFRASTRUCTURE ONLY — numpy-facing wrapper over the plain-C oracle (oracle/r3oracle.c).

The oracle is the CPU restatement of the reference algorithms used as the parity checker in
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product package
(r3det-pytorch_b200/) never imports it.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libr3oracle.so")
_lib = None

V1, V2, V3 = 1, 2, 3
_VARIANT = {"v1": 1, "v2": 2, "v3": 3, 1: 1, 2: 2, 3: 3}
_MODE = {"iou": 0, "iof": 1}
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH) or os.path.getmtime(_PATH) < os.path.getmtime(os.path.join(_HERE, "r3oracle.c")):
            from . import build
            if not build.main():
                raise RuntimeError("oracle build failed")
        _lib = C.CDLL(_PATH)
        _lib.orc_iou_matrix_f32.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        _lib.orc_iou_matrix_f32.restype = None
        _lib.orc_iou_aligned_f32.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, C.c_int, C.c_int, _f32p]
        _lib.orc_iou_aligned_f32.restype = None
        _lib.orc_iou_exact_f64.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, C.c_int, _f64p]
        _lib.orc_iou_exact_f64.restype = None
        _lib.orc_nms_f32.argtypes = [_f32p, C.c_int64, _f32p, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_int, C.c_int, _i64p]
        _lib.orc_nms_f32.restype = C.c_int64
        _lib.orc_poly_iou_aligned_f32.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int64, _f32p]
        _lib.orc_poly_iou_aligned_f32.restype = None
        _lib.orc_poly_nms_f32.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_float, _i64p]
        _lib.orc_poly_nms_f32.restype = C.c_int64
        _lib.orc_frm_forward_f32.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _f32p]
        _lib.orc_frm_forward_f32.restype = None
        _lib.orc_frm_backward_f32.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _f32p]
        _lib.orc_frm_backward_f32.restype = None
    return _lib


def _b5(a):
    a = np.asarray(a, dtype=np.float32)
    return np.ascontiguousarray(a.reshape(-1, a.shape[-1])[:, :5])


def iou_matrix(b1, b2, variant="v1", mode="iou", wrapper_mask=True, sort_kind=0):
    """(m,5) x (n,5) -> (m,n) float32 under the variant's reference semantics."""
    b1, b2 = _b5(b1), _b5(b2)
    out = np.zeros((b1.shape[0], b2.shape[0]), np.float32)
    if out.size:
        lib().orc_iou_matrix_f32(b1, b1.shape[0], b2, b2.shape[0], _VARIANT[variant], _MODE[mode],
                                 int(wrapper_mask), sort_kind, out)
    return out


def iou_aligned(b1, b2, variant="v1", mode="iou", sort_kind=0):
    b1, b2 = _b5(b1), _b5(b2)
    n = max(b1.shape[0], b2.shape[0])
    out = np.zeros((n,), np.float32)
    if b1.shape[0] and b2.shape[0]:
        lib().orc_iou_aligned_f32(b1, b1.shape[0], b2, b2.shape[0], _VARIANT[variant], _MODE[mode], sort_kind, out)
    return out


def iou_exact(b1, b2, variant="v1", mode="iou"):
    """float64 Sutherland-Hodgman ground truth (not a reference algorithm)."""
    b1, b2 = _b5(b1), _b5(b2)
    out = np.zeros((b1.shape[0], b2.shape[0]), np.float64)
    if out.size:
        lib().orc_iou_exact_f64(b1, b1.shape[0], b2, b2.shape[0], _MODE[mode], -1 if _VARIANT[variant] == 3 else 1, out)
    return out


def nms(boxes, scores, thr, variant="v1", labels=None, inclusive=False, sort_kind=0):
    """Greedy rotated NMS -> kept original indices in descending-score order.

    inclusive=True is the reference CPU rule (IoU >= thr), False the reference GPU rule (IoU > thr).
    labels: IoU forced to 0 across labels (ml_nms_rotated); None = single class.
    """
    boxes = np.ascontiguousarray(np.asarray(boxes, np.float32))
    scores = np.ascontiguousarray(np.asarray(scores, np.float32))
    K = boxes.shape[0]
    keep = np.empty((K,), np.int64)
    if K == 0:
        return keep
    lab = None
    labp = None
    if labels is not None:
        lab = np.ascontiguousarray(np.asarray(labels, np.float32))
        labp = lab.ctypes.data_as(C.c_void_p)
    n = lib().orc_nms_f32(boxes, boxes.shape[1], scores, labp, K, float(thr), _VARIANT[variant], int(inclusive), sort_kind, keep)
    return keep[:n].copy()


def frm_forward(feat, boxes, scale, points=1):
    feat = np.ascontiguousarray(feat, np.float32)
    boxes = np.ascontiguousarray(boxes, np.float32)
    N, Cc, H, W = feat.shape
    out = np.empty_like(feat)
    lib().orc_frm_forward_f32(feat, boxes.reshape(-1), N, Cc, H, W, float(scale), int(points), out)
    return out


def frm_backward(gout, boxes, scale, points=1, acc64=False):
    gout = np.ascontiguousarray(gout, np.float32)
    boxes = np.ascontiguousarray(boxes, np.float32)
    N, Cc, H, W = gout.shape
    gin = np.empty_like(gout)
    lib().orc_frm_backward_f32(gout, boxes.reshape(-1), N, Cc, H, W, float(scale), int(points), int(acc64), gin)
    return gin


def poly_iou_aligned(p, q):
    """IoU of 4-point polygons, row by row (poly_nms_cuda.cu:122-143): p, q (n, >=8) -> (n,)."""
    p = np.ascontiguousarray(np.asarray(p, np.float32)); q = np.ascontiguousarray(np.asarray(q, np.float32))
    out = np.empty((p.shape[0],), np.float32)
    if out.size:
        lib().orc_poly_iou_aligned_f32(p, p.shape[1], q, q.shape[1], p.shape[0], out)
    return out


def poly_nms(dets, thr):
    """Greedy polygon NMS on (K, 9) rows [x0, y0, ..., x3, y3, score] (rule IoU > thr): kept indices, score order."""
    dets = np.ascontiguousarray(np.asarray(dets, np.float32))
    K = dets.shape[0]
    keep = np.empty((K,), np.int64)
    if K == 0:
        return keep
    n = lib().orc_poly_nms_f32(dets, dets.shape[1], np.ascontiguousarray(dets[:, 8]), K, float(thr), keep)
    return keep[:n].copy()
