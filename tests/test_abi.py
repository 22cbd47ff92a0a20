"""The C-ABI library: builds, loads, exports every symbol include/r3geo.h declares, and validates arguments
(no compute calls — this file runs on the CPU-only box)."""
import ctypes as C
import os
import re

import pytest

from tests.util import ROOT


@pytest.fixture(scope="module")
def lib():
    import r3det_b200
    return r3det_b200._lib.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "r3geo.h")).read()
    return sorted(set(re.findall(r"\b(r3g_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    import r3det_b200
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/r3geo.h but not exported by libr3geo.so"
    assert sorted(r3det_b200._lib.exported_symbols()) == names


def test_version_and_workspace_queries(lib):
    assert lib.r3g_version() >= 100
    n = C.c_size_t(0)
    assert lib.r3g_iou_workspace_bytes(1000, 200000, C.byref(n)) == 0 and n.value >= (1000 + 200000) * 32
    assert lib.r3g_nms_workspace_bytes(8000, C.byref(n)) == 0 and n.value > 8000 * 64
    assert lib.r3g_frm_backward_workspace_bytes(8, 128, 128, 1, C.byref(n)) == 0 and n.value > 8 * 128 * 128 * 4 * 4


def test_argument_errors_are_reported(lib):
    n = C.c_size_t(0)
    assert lib.r3g_iou_workspace_bytes(-1, 5, C.byref(n)) < 0
    assert b"r3g_iou_workspace_bytes" in lib.r3g_last_error()
    assert lib.r3g_iou_matrix_f32(None, 4, 5, None, 4, 5, 7, 0, 0, None, None, 0, None) < 0
    assert b"variant" in lib.r3g_last_error()
    assert lib.r3g_iou_matrix_f32(None, 4, 5, None, 4, 5, 1, 3, 0, None, None, 0, None) < 0
    assert b"mode" in lib.r3g_last_error()
    assert lib.r3g_frm_forward_f32(None, None, 1, 1, 4, 4, 1.0, 3, None, None) < 0
    assert b"points" in lib.r3g_last_error()
    assert lib.r3g_obb2poly_f32(None, 4, 9, None, None) < 0
    # empty problems are no-ops, not errors
    assert lib.r3g_iou_matrix_f32(None, 0, 5, None, 7, 5, 1, 0, 0, None, None, 0, None) == 0
    assert lib.r3g_obb2poly_f32(None, 0, 1, None, None) == 0


def test_no_cpu_fallback():
    import torch
    import r3det_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r3det_b200.rbbox_iou(torch.zeros(2, 5), torch.zeros(3, 5))
    with pytest.raises(AssertionError):      # the reference's own `assert features.is_cuda` (feature_refine_module.py:20)
        r3det_b200.feature_refine(torch.zeros(1, 2, 4, 4), torch.zeros(16, 5), 0.125, 1)
    from r3det_b200.fr import frm_forward
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        frm_forward(torch.zeros(1, 2, 4, 4), torch.zeros(16, 5), 0.125, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r3det_b200.obb2poly(torch.zeros(2, 5), 'v1')


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "r3det-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+[<\"].*oracle", text, re.M), \
                    f"{f} pulls in the oracle"


def test_registry_names():
    import r3det_b200
    for name in ("RBboxOverlaps2D_v1", "RBboxOverlaps2D_v2", "RBboxOverlaps2D_v3"):
        calc = r3det_b200.IOU_CALCULATORS.build(dict(type=name))
        assert repr(calc) == name + "()"


def test_next_row_entry_points_validate_arguments(lib):
    """coder / dense-head tail / multi-level FRM / polygon NMS / assigner: size queries and argument errors (host only)."""
    f5 = (C.c_float * 5)(0, 0, 0, 0, 0); one5 = (C.c_float * 5)(1, 1, 1, 1, 1)
    # coder
    assert lib.r3g_delta2bbox_f32(None, 4, 5, None, 1, f5, one5, 4, None, 0.016, 0, 32.0, None, None) < 0
    assert b"variant" in lib.r3g_last_error()
    assert lib.r3g_delta2bbox_f32(None, 4, 5, None, 1, None, one5, 1, None, 0.016, 0, 32.0, None, None) < 0
    assert lib.r3g_delta2bbox_f32(None, 4, 5, None, 1, f5, one5, 1, None, -1.0, 0, 32.0, None, None) < 0
    assert b"wh_ratio_clip" in lib.r3g_last_error()
    assert lib.r3g_delta2bbox_f32(None, 0, 5, None, 1, f5, one5, 1, None, 0.016, 0, 32.0, None, None) == 0
    assert lib.r3g_bbox2delta_f32(None, 5, None, 5, 0, f5, one5, 3, None, None) == 0
    assert lib.r3g_bbox2delta_f32(None, 4, None, 5, 3, f5, one5, 3, None, None) < 0
    # get_bboxes tail: rows per image = sum over levels of min(nms_pre, H*W*A)
    hw = (C.c_int64 * 10)(128, 128, 64, 64, 32, 32, 16, 16, 8, 8)
    rows, nbytes = C.c_int64(0), C.c_size_t(0)
    assert lib.r3g_select_decode_sizes(5, 8, 9, hw, 2000, C.byref(rows), C.byref(nbytes)) == 0
    assert rows.value == 2000 * 4 + 8 * 8 * 9 == 8576 and nbytes.value > 8 * 196416 * 20
    assert lib.r3g_select_decode_sizes(5, 8, 9, hw, -1, C.byref(rows), C.byref(nbytes)) == 0 and rows.value == 196416
    assert lib.r3g_select_decode_sizes(9, 8, 9, hw, 2000, C.byref(rows), C.byref(nbytes)) < 0
    assert b"levels" in lib.r3g_last_error()
    assert lib.r3g_select_decode_sizes(5, 65, 9, hw, 2000, C.byref(rows), C.byref(nbytes)) < 0
    assert b"images" in lib.r3g_last_error()
    # FRM over several levels
    ihw = (C.c_int * 4)(16, 16, 8, 8)
    assert lib.r3g_frm_backward_multi_workspace_bytes(2, 2, ihw, 5, C.byref(nbytes)) == 0 and nbytes.value > 2 * 320 * 20 * 4
    assert lib.r3g_frm_backward_multi_workspace_bytes(9, 2, ihw, 5, C.byref(nbytes)) < 0
    sc = (C.c_float * 2)(0.125, 0.0625)
    assert lib.r3g_frm_forward_multi_f32(2, None, None, None, 2, 4, ihw, sc, 1, None, None) < 0      # null level pointers
    assert lib.r3g_frm_forward_multi_f32(2, None, None, None, 0, 4, ihw, sc, 1, None, None) == 0     # empty batch
    assert lib.r3g_frm_forward_multi_f32(2, None, None, None, 2, 4, ihw, sc, 2, None, None) < 0
    assert b"points" in lib.r3g_last_error()
    # polygon NMS
    assert lib.r3g_poly_nms_workspace_bytes(5000, C.byref(nbytes)) == 0 and nbytes.value > 5000 * 48
    keep = C.c_int64(0)
    assert lib.r3g_poly_nms_f32(None, 9, None, None, 5, 0.1, None, None, None, 0, None) < 0         # null num_keep_out
    # assigner
    assert lib.r3g_assign_workspace_bytes(1000, 200000, C.byref(nbytes)) == 0 and nbytes.value > 200000 * 12
    assert lib.r3g_max_iou_assign_f32(None, 3, 5, None, 4, 5, 9, 0, 0.5, 0.4, 0.0, 1, 1, None, None, None, None, None, None, 0, None) < 0
