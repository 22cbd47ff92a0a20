"""Test-only: host (g++) build of the product geometry headers, for GPU-less parity sweeps."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "r3det-pytorch_b200", "csrc")
_OUT = os.path.join(_HERE, "_build", "libhostgeom.so")
_lib = None
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_V = {"v1": 1, "v2": 2, "v3": 3}
_M = {"iou": 0, "iof": 1}


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "hostgeom.cpp"), os.path.join(_CSRC, "geom.cuh"), os.path.join(_CSRC, "emu.cuh")]
        if not os.path.exists(_OUT) or os.path.getmtime(_OUT) < max(os.path.getmtime(s) for s in srcs):
            os.makedirs(os.path.dirname(_OUT), exist_ok=True)
            # -ffp-contract=off + -O2: same FP behaviour as nvcc's non-contracted intrinsics in emu.cuh
            subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                            f"-I{_CSRC}", srcs[0], "-o", _OUT], check=True)
        _lib = C.CDLL(_OUT)
        _lib.hg_iou_matrix.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, _f32p,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib.hg_emu_matrix.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, C.c_int, _f32p]
    return _lib


def iou_matrix(b1, b2, variant="v1", mode="iou", tau=2e-2, stats=None, refit=True):
    b1 = np.ascontiguousarray(np.asarray(b1, np.float32)[:, :5])
    b2 = np.ascontiguousarray(np.asarray(b2, np.float32)[:, :5])
    out = np.zeros((len(b1), len(b2)), np.float32)
    nc, ns, ne = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    lib().hg_iou_matrix(b1, len(b1), b2, len(b2), _V[variant], _M[mode], tau, int(refit), out, C.byref(nc), C.byref(ns), C.byref(ne))
    if stats is not None:
        stats.update(circle_pass=nc.value, sat_pass=ns.value, emu=ne.value)
    return out


def emu_matrix(b1, b2, variant="v1", mode="iou"):
    b1 = np.ascontiguousarray(np.asarray(b1, np.float32)[:, :5])
    b2 = np.ascontiguousarray(np.asarray(b2, np.float32)[:, :5])
    out = np.zeros((len(b1), len(b2)), np.float32)
    lib().hg_emu_matrix(b1, len(b1), b2, len(b2), _V[variant], _M[mode], out)
    return out
