"""ctypes loader for libr3geo.so (the C ABI in include/r3geo.h) + tensor plumbing.

There is NO CPU fallback anywhere in this package: if the CUDA library has not been built
(`python r3det-pytorch_b200/build.py`) every op raises; CPU tensors raise.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("R3G_LIB", os.path.join(_HERE, "libr3geo.so"))      # R3G_LIB: an alternative build (variant timing)

V = {"v1": 1, "v2": 2, "v3": 3}
MODE = {"iou": 0, "iof": 1}
FLAG_STRICT = 1
FLAG_SMALL_MASK = 2
FLAG_EMULATE_ALL = 4
NMS_INCLUSIVE = 1
NMS_ORDER_INDEX = 2
NMS_DROP_SMALL = 4
NMS_STRICT = 8
NMS_SORT_PATH = 16

_lib = None

_vp, _i64, _i32, _f32, _sz, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t, C.c_double

_SIGNATURES = {
    "r3g_last_error": (C.c_char_p, []),
    "r3g_version": (_i32, []),
    "r3g_iou_workspace_bytes": (_i32, [_i64, _i64, C.POINTER(_sz)]),
    "r3g_iou_matrix_f32": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "r3g_iou_prepare_f32": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _vp, _sz, _vp]),
    "r3g_iou_matrix_prepared_f32": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "r3g_iou_aligned_f32": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp]),
    "r3g_assign_workspace_bytes": (_i32, [_i64, _i64, C.POINTER(_sz)]),
    "r3g_max_iou_assign_f32": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i32, _i32, _f32, _f32, _f32, _i32, _i32,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "r3g_assign_batched_workspace_bytes": (_i32, [_i64, _vp, _i64, _i32, C.POINTER(_sz)]),
    "r3g_max_iou_assign_batched_f32": (_i32, [_i64, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _f32, _f32, _i32, _i32,
                                              _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "r3g_nms_workspace_bytes": (_i32, [_i64, C.POINTER(_sz)]),
    "r3g_nms_f32": (_i32, [_vp, _i64, _vp, _vp, _i64, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "r3g_nms_batched_f32": (_i32, [_vp, _i64, _vp, _vp, _vp, _i32, _i64, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "r3g_mc_candidates_batched_f32": (_i32, [_vp, _i32, _vp, _i64, _i64, _i32, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                             _vp, _sz, _vp]),
    "r3g_nms_pack_f32": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "r3g_nms_batched_counted_f32": (_i32, [_vp, _i64, _vp, _vp, _vp, _i32, _i64, _vp, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "r3g_poly_nms_workspace_bytes": (_i32, [_i64, C.POINTER(_sz)]),
    "r3g_poly_nms_f32": (_i32, [_vp, _i64, _vp, _vp, _i64, _f32, _vp, _vp, _vp, _sz, _vp]),
    "r3g_mc_candidates_workspace_bytes": (_i32, [_i64, _i32, C.POINTER(_sz)]),
    "r3g_mc_candidates_f32": (_i32, [_vp, _i32, _vp, _i64, _vp, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "r3g_frm_forward_f32": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp]),
    "r3g_frm_backward_workspace_bytes": (_i32, [_i32, _i32, _i32, _i32, C.POINTER(_sz)]),
    "r3g_frm_backward_f32": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp, _sz, _vp]),
    "r3g_frm_forward_multi_f32": (_i32, [_i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "r3g_frm_backward_multi_workspace_bytes": (_i32, [_i32, _i32, _vp, _i32, C.POINTER(_sz)]),
    "r3g_frm_backward_multi_f32": (_i32, [_i32, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "r3g_frm_backward_plan_multi_f32": (_i32, [_i32, _vp, _i32, _vp, _vp, _i32, _vp, _sz, _vp]),
    "r3g_frm_backward_apply_multi_f32": (_i32, [_i32, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "r3g_obb2poly_f32": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "r3g_poly2obb_f32": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "r3g_obb2hbb_f32": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "r3g_hbb2obb_f32": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "r3g_obb2xyxy_f32": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "r3g_delta2bbox_f32": (_i32, [_vp, _i64, _i64, _vp, _i64, _vp, _vp, _i32, _vp, _f64, _i32, _f32, _vp, _vp]),
    "r3g_bbox2delta_f32": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp]),
    "r3g_filter_bboxes_f32": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _i32, _f64, _i32, _f32, _vp, _vp]),
    "r3g_refine_bboxes_f32": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _i32, _f64, _i32, _f32, _vp, _vp]),
    "r3g_select_decode_sizes": (_i32, [_i64, _i64, _i64, _vp, _i64, C.POINTER(_i64), C.POINTER(_sz)]),
    "r3g_select_decode_f32": (_i32, [_i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _i32, _f64, _i32, _f32,
                                     _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}


def exported_symbols():
    """Every symbol include/r3geo.h declares (used by the CPU-side ABI test)."""
    return list(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python r3det-pytorch_b200/build.py` "
                "(there is no CPU or PyTorch fallback for these ops)")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)   # AttributeError here = the .so is stale: rebuild
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"libr3geo error {rc}: {lib().r3g_last_error().decode()}")


def require_cuda(*tensors):
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"expected a torch.Tensor, got {type(t)}")
        if not t.is_cuda:
            raise RuntimeError("r3det_b200 ops run on CUDA tensors only (no CPU fallback); got a CPU tensor")


def require_no_grad(op, *tensors):
    """The kernels behind `op` write into fresh outputs and carry no autograd graph.  Where the reference's implementation is
    differentiable torch code, silently returning a detached result would zero a gradient: raise instead."""
    if torch.is_grad_enabled():
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.requires_grad:
                raise RuntimeError(f"r3det_b200.{op}: input requires grad, but this op is not differentiable here "
                                   "(detach the input, or call it under torch.no_grad())")


class device_guard:
    """`with device_guard(dev):` — like torch.cuda.device(dev) but free when dev is already current."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        self.ctx = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else C.c_void_p(0)


def as_f32_rows(t, min_cols=5):
    """float32, contiguous rows; returns (tensor, row stride in floats)."""
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    return t, (t.size(-1) if t.dim() > 1 else min_cols)


def workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
