"""N>1 host logic of the path on CPU: world_size-2 gloo processes exercise the row-sharded IoU statistics and
the single all_gather of keep lists (r3det-pytorch_b200/sharding.py).  The compute callable is injected — here
the oracle stands in for the CUDA op (tests only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import ROOT, clustered, rand_obb


def _worker(rank, world, port_no, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import r3det_b200  # noqa: F401
    from r3det_b200 import sharding
    from oracle import port
    from collections import namedtuple
    Out = namedtuple("Out", "max_overlaps gt_max_overlaps gt_argmax_overlaps")

    def assign_fn(g, a):          # stands in for r3det_b200.max_iou_assign (the fused CUDA assigner): same three outputs
        m = torch.from_numpy(port.iou_matrix(g.numpy(), a.numpy(), "v1"))
        return Out(m.max(dim=0).values, m.max(dim=1).values, m.max(dim=1).indices)

    gt = torch.from_numpy(rand_obb(40, 1)); anchors = torch.from_numpy(rand_obb(1001, 2))
    out, lo, hi = sharding.sharded_assign(gt, anchors, assign_fn)
    gmax, garg, npos, nneg = sharding.assigner_stats(out.gt_max_overlaps, out.gt_argmax_overlaps, out.max_overlaps, lo, 0.5, 0.4)
    # image-sharded keep lists
    num_images, max_per = 5, 30
    ilo, ihi = sharding.shard_range(num_images, rank, world)
    dets, labels = [], []
    for i in range(ilo, ihi):
        b, s, l = clustered(20 + 7 * i, 100 + i)
        keep = port.nms(b, s, 0.1, "v1")[:max_per]
        dets.append(torch.from_numpy(np.concatenate([b[keep], s[keep, None]], 1)))
        labels.append(torch.from_numpy(l[keep]))
    all_dets, all_labels = sharding.gather_keep_lists(dets, labels, max_per, num_images)
    # the asynchronous form and the padded-record form publish the same lists
    h = sharding.gather_keep_lists(dets, labels, max_per, num_images, async_op=True)
    d2, l2 = h.wait()
    per_rank = (num_images + world - 1) // world
    pd = torch.zeros((per_rank, max_per, 6)); pl = torch.zeros((per_rank, max_per), dtype=torch.int64); pc = torch.zeros((per_rank,), dtype=torch.int64)
    for i, (d, l) in enumerate(zip(dets, labels)):
        pd[i, :d.size(0)] = d; pl[i, :l.size(0)] = l; pc[i] = d.size(0)
    d3, l3 = sharding.gather_padded_records(pd, pl, pc, num_images)
    d4, l4, c4 = sharding.gather_padded_records(pd, pl, pc, num_images, padded=True)
    assert d4.shape == (num_images, max_per, 6) and c4.tolist() == [d.size(0) for d in all_dets]
    for i, d in enumerate(all_dets):
        assert torch.equal(d4[i, :d.size(0)], d) and torch.equal(l4[i, :d.size(0)], all_labels[i])
    g2 = sharding.assigner_stats(out.gt_max_overlaps, out.gt_argmax_overlaps, out.max_overlaps, lo, 0.5, 0.4, sync=False)
    assert int(g2[2]) == npos and int(g2[3]) == nneg and torch.equal(g2[0], gmax)
    for a, b_, c in zip(all_dets, d2, d3):
        assert torch.equal(a, b_) and torch.equal(a, c)
    for a, b_, c in zip(all_labels, l2, l3):
        assert torch.equal(a, b_) and torch.equal(a, c)
    q.put((rank, lo, hi, gmax.numpy(), garg.numpy(), npos, nneg, [d.numpy() for d in all_dets], [l.numpy() for l in all_labels]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo():
    from oracle import port
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # unsharded truth
    gt, anchors = rand_obb(40, 1), rand_obb(1001, 2)
    full = port.iou_matrix(gt, anchors, "v1")
    amax = full.max(0)
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 501, 501, 1001)
    for r in res:
        assert np.array_equal(r[3], full.max(1))
        assert np.array_equal(r[4], full.argmax(1))
        assert r[5] == int((amax >= 0.5).sum()) and r[6] == int(((amax >= 0) & (amax < 0.4)).sum())
    for i in range(5):
        b, s, l = clustered(20 + 7 * i, 100 + i)
        keep = port.nms(b, s, 0.1, "v1")[:30]
        want = np.concatenate([b[keep], s[keep, None]], 1)
        for r in res:
            assert np.array_equal(r[7][i], want) and np.array_equal(r[8][i], l[keep])


def test_shard_range_partitions():
    from r3det_b200 import sharding
    for n in (0, 1, 7, 64, 200000):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
