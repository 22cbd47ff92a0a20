"""The library never allocates, synchronises or reads device results on the host, so a sequence of its calls can be captured
in a CUDA graph and replayed on new data (launch-bound inner loops: per-image assignment + FRM + NMS)."""
import numpy as np
import pytest
import torch

from tests.util import clustered, rand_obb

pytestmark = pytest.mark.gpu


def test_cuda_graph_capture_and_replay(cuda_dev):
    import r3det_b200 as R
    from r3det_b200._nms_core import nms_device
    from r3det_b200.fr import frm_backward_multi, frm_forward_multi
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda_dev)
    gt, an = t(rand_obb(50, 1, "v1")), t(rand_obb(20000, 2, "v1"))
    b, s, l = clustered(3000, 3, "v1")
    B, S, Lb = t(b), t(s), t(l)
    feats = [torch.randn(2, 16, h, h, device=cuda_dev) for h in (32, 16)]
    boxes = [t(np.concatenate([np.random.default_rng(h).uniform(0, 256, (2 * h * h, 2)), np.random.default_rng(h).uniform(8, 64, (2 * h * h, 2)),
                               np.random.default_rng(h).uniform(-1.5, 0, (2 * h * h, 1))], 1).astype(np.float32)) for h in (32, 16)]
    scale = torch.tensor(1100.0, device=cuda_dev)

    def work():
        iou = R.pairwise_iou(gt, an, "v1")
        asg = R.max_iou_assign(gt, an, 0.5, 0.4, 0.0, True, True, "v1")
        y = frm_forward_multi(feats, boxes, [1 / 8, 1 / 16], 5)
        g = frm_backward_multi(feats, boxes, [1 / 8, 1 / 16], 5)
        keep, num = nms_device(B, S, 0.1, "v1", labels=Lb, class_offset=scale, order_index=True)
        return iou, asg.gt_inds, asg.max_overlaps, y[0], y[1], g[0], g[1], keep, num

    side = torch.cuda.Stream(device=cuda_dev)
    side.wait_stream(torch.cuda.current_stream(cuda_dev))
    with torch.cuda.stream(side):
        for _ in range(2):
            work()                                   # warm-up outside the capture (attribute setup, allocator pools)
    torch.cuda.current_stream(cuda_dev).wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = work()
    # new data in the captured input buffers, replay, compare with an eager run on the same data
    an.copy_(t(rand_obb(20000, 7, "v1")))
    b2, s2, _ = clustered(3000, 9, "v1")
    B.copy_(t(b2)); S.copy_(t(s2))
    for f in feats:
        f.normal_()
    graph.replay()
    torch.cuda.synchronize()
    got = [o.clone() for o in outs]
    want = work()
    torch.cuda.synchronize()
    n = int(want[-1])
    assert int(got[-1]) == n and torch.equal(got[-2][:n], want[-2][:n])
    for a_, b_ in zip(got[:-2], want[:-2]):
        assert torch.equal(a_, b_)
