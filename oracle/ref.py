"""TEST INFRASTRUCTURE ONLY — ctypes access to the unmodified reference CPU ops in oracle/_ref/.

Built by oracle/build_ref.py from the sources under /root/reference (never copied here).
Used to pin the C restatement in oracle/r3oracle.c, to generate tests/golden/, and as the
``cpu_baseline.kind == "reference"`` / ``bench.py --impl reference`` timing arm.
The product package must never import this module.
"""
import ctypes as C
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_libs = {}

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")


def available(name="libref_v1.so"):
    return os.path.exists(os.path.join(_DIR, name))


def _lib(name):
    if name not in _libs:
        path = os.path.join(_DIR, name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
        if name != "libref_v2.so":
            import torch  # noqa: F401  (libtorch must be resolvable; rpath handles the rest)
        _libs[name] = C.CDLL(path)
    return _libs[name]


def _as(a, dt):
    return np.ascontiguousarray(np.asarray(a), dtype=dt)


def v1_iou(b1, b2, iof=False, dtype=np.float32):
    """rbbox_geo mat_iou_iof semantics driven over the reference's host geometry templates."""
    b1, b2 = _as(b1, dtype)[:, :5].copy(), _as(b2, dtype)[:, :5].copy()
    out = np.empty((b1.shape[0], b2.shape[0]), dtype)
    lib = _lib("libref_v1.so")
    if dtype == np.float32:
        f = lib.ref_v1_iou_matrix_f32
        f.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, _f32p]
    else:
        f = lib.ref_v1_iou_matrix_f64
        f.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, C.c_int, _f64p]
    f.restype = None
    f(b1, b1.shape[0], b2, b2.shape[0], int(iof), out)
    return out


def v1_iou_aligned(b1, b2, iof=False):
    b1, b2 = _as(b1, np.float32)[:, :5].copy(), _as(b2, np.float32)[:, :5].copy()
    out = np.empty((max(b1.shape[0], b2.shape[0]),), np.float32)
    f = _lib("libref_v1.so").ref_v1_iou_aligned_f32
    f.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, _f32p]
    f.restype = None
    f(b1, b1.shape[0], b2, b2.shape[0], int(iof), out)
    return out


def v1_nms(dets6, thr):
    """rnms_ext.rnms on CPU: dets (K,6) -> ascending kept original indices (uses >=)."""
    d = _as(dets6, np.float32)
    keep = np.empty((d.shape[0],), np.int64)
    if d.shape[0] == 0:
        return keep
    f = _lib("libref_v1.so").ref_v1_nms_f32
    f.argtypes = [_f32p, C.c_int64, C.c_float, _i64p]
    f.restype = C.c_int64
    n = f(d, d.shape[0], float(thr), keep)
    return keep[:n].copy()


def v3_iou(b1, b2, mode_iou=True, dtype=np.float32, through_tensor=False):
    """box_iou_rotated_ext.overlaps on CPU (flag True == IoU)."""
    b1, b2 = _as(b1, dtype)[:, :5].copy(), _as(b2, dtype)[:, :5].copy()
    out = np.empty((b1.shape[0], b2.shape[0]), dtype)
    lib = _lib("libref_v3iou.so")
    if dtype == np.float32:
        f = lib.ref_v3_iou_matrix_tensor_f32 if through_tensor else lib.ref_v3_iou_matrix_f32
        f.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, _f32p]
    else:
        f = lib.ref_v3_iou_matrix_f64
        f.argtypes = [_f64p, C.c_int64, _f64p, C.c_int64, C.c_int, _f64p]
    f.restype = None
    f(b1, b1.shape[0], b2, b2.shape[0], int(mode_iou), out)
    return out


def v3_nms(dets5, scores, thr):
    """nms_rotated_ext.nms_rotated on CPU: keep in descending-score order (uses >=)."""
    d, s = _as(dets5, np.float32), _as(scores, np.float32)
    keep = np.empty((d.shape[0],), np.int64)
    if d.shape[0] == 0:
        return keep
    f = _lib("libref_v3nms.so").ref_v3_nms_f32
    f.argtypes = [_f32p, _f32p, C.c_int64, C.c_float, _i64p]
    f.restype = C.c_int64
    n = f(d, s, d.shape[0], float(thr), keep)
    return keep[:n].copy()


def v2_iou(b1, b2, labels1=None, labels2=None):
    """ml_nms_rotated header: IoU with rotation +a, 0 across different labels."""
    b1, b2 = _as(b1, np.float32)[:, :5], _as(b2, np.float32)[:, :5]
    l1 = np.zeros(len(b1), np.float32) if labels1 is None else _as(labels1, np.float32)
    l2 = np.zeros(len(b2), np.float32) if labels2 is None else _as(labels2, np.float32)
    a = np.ascontiguousarray(np.concatenate([b1, l1[:, None]], 1))
    b = np.ascontiguousarray(np.concatenate([b2, l2[:, None]], 1))
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    f = _lib("libref_v2.so").ref_v2_iou_matrix_f32
    f.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, _f32p]
    f.restype = None
    f(a, a.shape[0], b, b.shape[0], out)
    return out


def v2_nms(dets5, scores, labels, thr):
    d = _as(dets5, np.float32)
    d6 = np.ascontiguousarray(np.concatenate([d, _as(labels, np.float32)[:, None]], 1))
    s = _as(scores, np.float32)
    keep = np.empty((d.shape[0],), np.int64)
    if d.shape[0] == 0:
        return keep
    f = _lib("libref_v2.so").ref_v2_nms_f32
    f.argtypes = [_f32p, _f32p, C.c_int64, C.c_float, _i64p]
    f.restype = C.c_int64
    n = f(d6, s, d.shape[0], float(thr), keep)
    return keep[:n].copy()
