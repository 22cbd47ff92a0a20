"""The oracle (oracle/r3oracle.c, oracle/transforms_np.py) against the committed golden vectors, which were
produced by the reference's own binaries and Python (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from tests.util import golden
from oracle import port
from oracle import transforms_np as tnp


def test_iou_v1_bitexact():
    g = golden("iou_ref.npz")
    assert np.array_equal(port.iou_matrix(g["v1_b1"], g["v1_b2"], "v1"), g["v1_iou"])
    assert np.array_equal(port.iou_matrix(g["v1_b1"], g["v1_b2"], "v1", "iof"), g["v1_iof"])
    assert np.array_equal(port.iou_aligned(g["v1_b1"][:100], g["v1_b2"][:100], "v1"), g["v1_aligned"])
    assert np.array_equal(port.iou_aligned(g["v1_b1"][:1], g["v1_b2"][:50], "v1"), g["v1_aligned_bcast"])


def test_iou_v3_bitexact_host_sort_and_device_sort():
    g = golden("iou_ref.npz")
    host = port.iou_matrix(g["v3_b1"], g["v3_b2"], "v3", wrapper_mask=False, sort_kind=1)
    assert np.array_equal(host, g["v3_iou"])
    assert np.array_equal(port.iou_matrix(g["v3_b1"], g["v3_b2"], "v3", "iof", wrapper_mask=False, sort_kind=1), g["v3_iof"])
    dev = port.iou_matrix(g["v3_b1"], g["v3_b2"], "v3", wrapper_mask=False, sort_kind=0)   # the __CUDACC__ sort branch
    assert np.abs(dev - g["v3_iou"]).max() <= 1e-6


def test_iou_v2_bitexact():
    g = golden("iou_ref.npz")
    assert np.array_equal(port.iou_matrix(g["v2_b1"], g["v2_b2"], "v2", sort_kind=1), g["v2_iou"])


def test_known_answer_pair():
    """(50,50,40,20,0.3) vs (55,48,30,25,-0.5) -> IoU 0.545840204 / IoF 0.684136271 (v3)."""
    g = golden("iou_ref.npz")
    iou = port.iou_matrix(g["kat_b1"], g["kat_b2"], "v3")[0, 0]
    iof = port.iou_matrix(g["kat_b1"], g["kat_b2"], "v3", "iof")[0, 0]
    assert abs(iou - 0.545840204) < 1e-6 and abs(iof - 0.684136271) < 1e-6
    assert iou == g["kat_v3_iou"][0, 0] and iof == g["kat_v3_iof"][0, 0]
    neg = lambda b: b * np.array([1, 1, 1, 1, -1], np.float32)   # v3 rotates by -a: v1(-a) == v3(a)
    assert abs(port.iou_matrix(neg(g["kat_b1"]), neg(g["kat_b2"]), "v1")[0, 0] - iou) < 1e-6
    assert abs(port.iou_exact(g["kat_b1"], g["kat_b2"], "v3")[0, 0] - iou) < 1e-6


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_nms_keep_sets(v):
    g = golden("nms_ref.npz")
    b, s, l = g[f"{v}_boxes"], g[f"{v}_scores"], g[f"{v}_labels"]
    if v == "v1":      # rnms_ext on CPU: >=, ascending original index
        k = np.sort(port.nms(b, s, 0.1, "v1", inclusive=True))
    elif v == "v3":    # nms_rotated_ext on CPU: >=, score order
        k = port.nms(b, s, 0.1, "v3", inclusive=True, sort_kind=1)
    else:              # ml_nms_rotated greedy loop with labels
        k = port.nms(b, s, 0.1, "v2", labels=l.astype(np.float32), inclusive=True, sort_kind=1)
    assert np.array_equal(k, g[f"{v}_keep"])


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_transforms(v):
    g = golden("transforms_ref.npz")
    obb = g[f"{v}_obb"]
    tol = dict(rtol=0, atol=2e-4)       # coordinates up to ~1300 px in float32 (ulp 1.2e-4), torch vs numpy trig
    assert np.allclose(tnp.obb2poly(obb, v), g[f"{v}_obb2poly"], **tol)
    assert np.allclose(tnp.obb2hbb(obb, v), g[f"{v}_obb2hbb"], **tol)
    assert np.allclose(tnp.obb2xyxy(obb, v), g[f"{v}_obb2xyxy"], **tol)
    assert np.allclose(tnp.hbb2obb(g[f"{v}_hbb"], v), g[f"{v}_hbb2obb"], **tol)
    for src, dst in ((f"{v}_obb2poly", f"{v}_poly2obb"), (f"{v}_rpoly", f"{v}_rpoly2obb")):
        got, want = tnp.poly2obb(g[src], v), g[dst]
        assert np.allclose(got[:, :4], want[:, :4], **tol)
        period = np.pi / 2 if v == "v1" else np.pi
        da = np.abs(got[:, 4] - want[:, 4]); da = np.minimum(da, np.abs(period - da))
        assert da.max() < 1e-4


def test_polygon_iou_and_nms_vs_reference_cuda():
    """poly_nms_cuda.cu compiled unmodified and run on a B200 (tests/golden/make_golden_gpu.py): the C restatement agrees
    to FP32 cancellation noise on the IoU values (the reference build contracts into FMAs) and exactly on the keep lists."""
    g = golden("poly_refcuda.npz")
    for tag in ("rect", "quad"):
        d = g[f"{tag}_dets"]
        pr = g[f"{tag}_pairs"]
        got = port.poly_iou_aligned(d[pr[:, 0]], d[pr[:, 1]])
        # the fans are anchored at the image origin: |coordinate| ~ 1e3 px turns FMA-vs-no-FMA rounding into ~1e-3 of IoU
        assert np.abs(got - g[f"{tag}_iou_pairs"]).max() < 2e-3 and np.abs(got - g[f"{tag}_iou_pairs"]).mean() < 1e-4
        assert (g[f"{tag}_iou_pairs"] > 0.05).sum() > 100                  # the fixture does exercise overlapping pairs
        assert np.array_equal(port.poly_nms(d, float(g[f"{tag}_thr"])), g[f"{tag}_keep"])


def test_v1_nms_gpu_rule_vs_reference_cuda():
    """rnms_kernel.cu compiled unmodified and run on a B200: rule IoU > thr, keep list in ascending index."""
    g = golden("poly_refcuda.npz")
    d = g["v1nms_dets"]
    assert np.array_equal(np.sort(port.nms(d[:, :5], d[:, 5], 0.1, "v1", inclusive=False)), g["v1nms_keep"])
