"""v1 rotated IoU — host-side mirror of r3det/ops/rbbox_geo/rbbox_geo.py:4-9.

``rbbox_iou(rb1, rb2, vec=False, iof=False)`` has the reference's signature and results:
matrix mode returns (m, n); vec mode returns (max(m, n),) with modulo broadcast
(rbbox_geo_kernel.cu:270-309).  CUDA tensors only, as in the reference (rbbox_geo_cuda.cpp:6-11).
"""
import ctypes as C

import torch

from . import _lib as L


def pairwise_iou(b1, b2, variant, mode="iou", flags=L.FLAG_STRICT, return_stats=False):
    """(m,>=5) x (n,>=5) CUDA float tensors -> (m, n) IoU/IoF under `variant` semantics."""
    L.require_cuda(b1, b2)
    b1, s1 = L.as_f32_rows(b1)
    b2, s2 = L.as_f32_rows(b2)
    m, n = b1.size(0), b2.size(0)
    out = torch.empty((m, n), dtype=torch.float32, device=b1.device)
    if m == 0 or n == 0:
        return (out, None) if return_stats else out
    lib = L.lib()
    nbytes = C.c_size_t(0)
    L.check(lib.r3g_iou_workspace_bytes(m, n, C.byref(nbytes)))
    ws = L.workspace(nbytes.value, b1.device)
    with L.device_guard(b1.device):
        L.check(lib.r3g_iou_matrix_f32(L.ptr(b1), m, s1, L.ptr(b2), n, s2, L.V[variant], L.MODE[mode], flags,
                                       L.ptr(out), L.ptr(ws), ws.numel(), L.stream_ptr(b1.device)))
    if return_stats:
        return out, ws[:32].view(torch.int64)
    return out


def aligned_iou(b1, b2, variant, mode="iou", flags=L.FLAG_STRICT):
    """element i = overlap(b1[i % n1], b2[i % n2]); length max(n1, n2)."""
    L.require_cuda(b1, b2)
    b1, s1 = L.as_f32_rows(b1)
    b2, s2 = L.as_f32_rows(b2)
    n1, n2 = b1.size(0), b2.size(0)
    out = torch.empty((max(n1, n2),), dtype=torch.float32, device=b1.device)
    if n1 == 0 or n2 == 0:
        return out
    with L.device_guard(b1.device):
        L.check(L.lib().r3g_iou_aligned_f32(L.ptr(b1), n1, s1, L.ptr(b2), n2, s2, L.V[variant], L.MODE[mode],
                                            flags, L.ptr(out), L.stream_ptr(b1.device)))
    return out


def rbbox_iou(rb1, rb2, vec=False, iof=False):
    """Compute the IoU of oriented bboxes (reference: r3det/ops/rbbox_geo/rbbox_geo.py:4-9)."""
    mode = "iof" if iof else "iou"
    out = aligned_iou(rb1, rb2, "v1", mode) if vec else pairwise_iou(rb1, rb2, "v1", mode)
    # the reference dispatches on the input dtype (rbbox_geo_kernel.cu:318) and returns it; arithmetic here is FP32
    return out if rb1.dtype == torch.float32 or not rb1.is_floating_point() else out.to(rb1.dtype)
