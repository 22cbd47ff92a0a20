"""Robustness of the C-ABI path on the GPU: caller streams, non-contiguous / strided inputs, NaN and degenerate boxes
(no hang, no crash, sane outputs), repeated calls on one workspace."""
import numpy as np
import pytest
import torch

from tests.util import clustered, rand_obb
from oracle import port

pytestmark = pytest.mark.gpu


def _t(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def test_caller_stream_and_strided_inputs(cuda_dev):
    import r3det_b200 as R
    a, b = rand_obb(300, 1), rand_obb(2000, 2)
    want = port.iou_matrix(a, b, "v1")
    side = torch.cuda.Stream(device=cuda_dev)
    A6 = torch.cat([_t(a, cuda_dev), torch.rand(300, 1, device=cuda_dev)], 1)       # row stride 6 via the calculator
    wide = torch.zeros((2000, 9), device=cuda_dev); wide[:, 2:7] = _t(b, cuda_dev)   # non-contiguous view
    with torch.cuda.stream(side):
        got = R.RBboxOverlaps2D_v1()(A6, wide[:, 2:7])
        keep = R.rnms(torch.cat([_t(b, cuda_dev), torch.rand(2000, 1, device=cuda_dev)], 1), 0.3)[1]
    side.synchronize()
    assert np.abs(got.cpu().numpy() - want).max() <= 1e-5
    assert keep.numel() > 0


def test_nan_inf_and_degenerate_boxes_do_not_hang(cuda_dev):
    import r3det_b200 as R
    b = rand_obb(500, 3)
    b[::50, 0] = np.nan; b[1::50, 2] = np.inf; b[2::50, 3] = 0.0; b[3::50, 2] = -5.0; b[4::50, 4] = 1e30
    B = _t(b, cuda_dev)
    for v in ("v1", "v2", "v3"):
        out = R.pairwise_iou(B, B, v)
        torch.cuda.synchronize()
        clean = np.ones(500, bool); clean[::50] = clean[1::50] = clean[2::50] = clean[3::50] = clean[4::50] = False
        sub = out.cpu().numpy()[np.ix_(clean, clean)]
        assert np.abs(sub - port.iou_matrix(b[clean], b[clean], v, wrapper_mask=False)).max() <= 1e-5   # clean pairs unaffected
        s = torch.rand(500, device=cuda_dev)
        from r3det_b200._nms_core import nms_device
        keep, num = nms_device(B, s, 0.3, v)
        assert 0 < int(num) <= 500


def test_repeated_calls_are_deterministic(cuda_dev):
    import r3det_b200 as R
    c, s, l = clustered(4000, 8, "v3")
    C, S, Lb = _t(c, cuda_dev), _t(s, cuda_dev), _t(l, cuda_dev)
    first = R.obb_batched_nms(C, S, Lb, 0.1)[1].cpu().numpy()
    for _ in range(5):
        assert np.array_equal(R.obb_batched_nms(C, S, Lb, 0.1)[1].cpu().numpy(), first)
    o1 = R.pairwise_iou(C[:500], C, "v3")
    assert torch.equal(o1, R.pairwise_iou(C[:500], C, "v3"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process():
    """One process driving two GPUs: per-device kernel attributes / occupancy caches, device guards, streams."""
    import r3det_b200 as R
    from r3det_b200._nms_core import nms_device
    from tests.util import clustered, rand_obb
    outs = []
    for idx in (0, 1, 0):
        dev = torch.device("cuda", idx)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        iou = R.pairwise_iou(t(rand_obb(70, 1)), t(rand_obb(5000, 2)), "v1")
        b, s, l = clustered(20000, 3, "v1")
        keep, num = nms_device(t(b), t(s), 0.1, "v1", labels=t(l))
        f = torch.from_numpy(np.random.default_rng(0).standard_normal((2, 8, 16, 16)).astype(np.float32)).to(dev)
        y = R.feature_refine(f, t(np.concatenate([np.random.default_rng(1).uniform(0, 128, (512, 2)), np.full((512, 2), 20.0),
                                                   np.zeros((512, 1))], 1).astype(np.float32)), 0.125, 5)
        assert iou.device == dev and keep.device == dev and y.device == dev
        outs.append((iou.cpu(), keep[:int(num)].cpu(), y.cpu()))
    for a, b_ in zip(outs[0], outs[1]):
        assert torch.equal(a, b_)
    for a, b_ in zip(outs[0], outs[2]):
        assert torch.equal(a, b_)


def test_half_and_double_inputs_keep_their_dtype(cuda_dev):
    """The reference ops dispatch on the input dtype and return it; here the arithmetic is FP32 and the result is cast."""
    import r3det_b200 as R
    a = torch.from_numpy(rand_obb(20, 1)).to(cuda_dev); b = torch.from_numpy(rand_obb(300, 2)).to(cuda_dev)
    for dt, tol in ((torch.float64, 1e-6), (torch.float16, 1e-3)):
        ref = R.rbbox_iou(a.to(dt).float(), b.to(dt).float())             # the same (dtype-rounded) boxes in FP32
        got = R.rbbox_iou(a.to(dt), b.to(dt))
        assert got.dtype == dt and (got.float() - ref).abs().max().item() <= tol
        got3 = R.obb_overlaps(a.to(dt), b.to(dt))
        assert got3.dtype == dt and got3.shape == (20, 300)
    calc = R.RBboxOverlaps2D_v1()
    assert calc(a.double(), b.double()).shape == (20, 300)
