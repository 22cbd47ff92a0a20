"""Fused max-IoU assignment vs the restatement of mmdet-2.19 MaxIoUAssigner (oracle/assign_np.py: provenance and known-answer
tests there; mmdet itself is unavailable -> parity unpinned upstream) applied to OUR overlap matrix: the fused kernel reduces
the very same values, so equality is exact, ties included."""
import numpy as np
import pytest
import torch

from tests.util import anchors_1024, rand_obb

pytestmark = pytest.mark.gpu


def assign_wrt_overlaps(overlaps, pos_iou_thr, neg_iou_thr, min_pos_iou, match_low_quality, gt_max_assign_all):
    """oracle/assign_np.py on the device matrix -> torch tensors on its device"""
    from oracle import assign_np
    got, mx = assign_np.assign_wrt_overlaps(overlaps.cpu().numpy(), pos_iou_thr, neg_iou_thr, min_pos_iou, match_low_quality,
                                            gt_max_assign_all)
    return torch.from_numpy(got).to(overlaps.device), torch.from_numpy(mx).to(overlaps.device)


@pytest.mark.parametrize("v", ["v1", "v3"])
@pytest.mark.parametrize("assign_all", [True, False])
def test_against_matrix_reduction(cuda_dev, v, assign_all):
    import r3det_b200 as R
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda_dev)
    cases = [(rand_obb(60, 1, v, 10, 300), anchors_1024(), 0.5, 0.4, 0.0),
             (rand_obb(200, 2, v), rand_obb(5003, 3, v), 0.6, 0.5, 0.3),
             (rand_obb(3, 4, v, 8, 16), rand_obb(1000, 5, v, 8, 16), 0.5, 0.4, 0.0)]      # GTs that overlap nothing
    # a symmetric set: exact ties between anchors
    g = np.array([[100, 100, 40, 40, 0.0]], np.float32)
    an = np.array([[90, 100, 40, 40, 0.0], [110, 100, 40, 40, 0.0], [100, 90, 40, 40, 0.0], [300, 300, 10, 10, 0.0]], np.float32)
    cases.append((g, an, 0.9, 0.4, 0.0))
    for gt, an, pos, neg, minpos in cases:
        G_, A_ = t(gt), t(an)
        if v == "v1":
            G_ = R.obb2hbb(G_, "v1") if gt.shape[0] > 3 else G_
        ov = R.pairwise_iou(G_, A_, v)
        want, want_max = assign_wrt_overlaps(ov, pos, neg, minpos, True, assign_all)
        out = R.max_iou_assign(G_, A_, pos, neg, minpos, True, assign_all, v)
        assert torch.equal(out.max_overlaps, want_max)
        assert torch.equal(out.argmax_overlaps, ov.max(dim=0)[1])
        assert torch.equal(out.gt_max_overlaps, ov.max(dim=1)[0]) and torch.equal(out.gt_argmax_overlaps, ov.max(dim=1)[1])
        assert torch.equal(out.gt_inds, want), (v, assign_all, gt.shape, int((out.gt_inds != want).sum()))
        nolq = R.max_iou_assign(G_, A_, pos, neg, minpos, False, assign_all, v)
        assert torch.equal(nolq.gt_inds, assign_wrt_overlaps(ov, pos, neg, minpos, False, assign_all)[0])


def test_edge_cases_and_class(cuda_dev):
    import r3det_b200 as R
    an = torch.from_numpy(rand_obb(500, 1)).to(cuda_dev)
    out = R.max_iou_assign(an[:0], an, 0.5, 0.4)
    assert out.num_gts == 0 and (out.gt_inds == 0).all() and (out.max_overlaps == 0).all()
    gt = torch.from_numpy(rand_obb(20, 2)).to(cuda_dev)
    labels = torch.arange(20, device=cuda_dev) % 15
    a = R.FusedMaxIoUAssigner(0.5, 0.4, min_pos_iou=0, iou_calculator=dict(type='RBboxOverlaps2D_v1'))
    res = a.assign(an, gt, gt_labels=labels)
    pos = res.gt_inds > 0
    assert pos.any() and torch.equal(res.labels[pos], labels[res.gt_inds[pos] - 1]) and (res.labels[~pos] == -1).all()


def test_adversarial_cases_vs_oracle(cuda_dev):
    """Zero rows with min_pos_iou = 0, duplicate anchors (exact ties), gt_max_assign_all False collisions, tuple neg_iou_thr,
    ignore regions (both ignore_wrt_candidates settings) — FusedMaxIoUAssigner vs oracle/assign_np.py on our own overlaps."""
    import r3det_b200 as R
    from oracle import assign_np
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda_dev)
    rng = np.random.default_rng(5)
    gt = rand_obb(12, 21, "v1", 20, 120)
    gt[3, :2] = [5000, 5000]                                              # a GT far outside: its row is all zeros
    an = rand_obb(3000, 22, "v1", 10, 150)
    an[100:140] = an[100]                                                 # 40 identical anchors: exact ties on every GT
    an[200:203] = gt[5]                                                   # anchors equal to a GT (IoU 1)
    an[300] = gt[6]; an[301] = gt[6]                                      # two perfect matches for one GT
    ign = rand_obb(4, 23, "v1", 100, 300)
    G_, A_, I_ = t(gt), t(an), t(ign)
    ov = R.pairwise_iou(G_, A_, "v1").cpu().numpy()
    for assign_all in (True, False):
        for mlq in (True, False):
            for neg in (0.4, (0.1, 0.4)):
                for minpos in (0.0, 0.3):
                    a = R.FusedMaxIoUAssigner(0.5, neg, min_pos_iou=minpos, gt_max_assign_all=assign_all, match_low_quality=mlq)
                    res = a.assign(A_, G_)
                    want, wmax = assign_np.assign_wrt_overlaps(ov, 0.5, neg, minpos, mlq, assign_all)
                    assert np.array_equal(res.gt_inds.cpu().numpy(), want), (assign_all, mlq, neg, minpos)
                    assert np.array_equal(res.max_overlaps.cpu().numpy(), wmax)
    for wrt in (True, False):
        iof = (R.pairwise_iou(A_, I_, "v1", "iof").max(dim=1).values if wrt else R.pairwise_iou(I_, A_, "v1", "iof").max(dim=0).values).cpu().numpy()
        assert (iof > 0.5).any() and not (iof > 0.5).all()
        a = R.FusedMaxIoUAssigner(0.5, 0.4, min_pos_iou=0.0, ignore_iof_thr=0.5, ignore_wrt_candidates=wrt)
        res = a.assign(A_, G_, gt_bboxes_ignore=I_)
        want, wmax = assign_np.assign(ov, iof, 0.5, pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0.0)
        assert np.array_equal(res.gt_inds.cpu().numpy(), want) and np.array_equal(res.max_overlaps.cpu().numpy(), wmax)
    # gpu_assign_thr is accepted (and has no effect: there is no CPU path)
    assert R.FusedMaxIoUAssigner(0.5, 0.4, gpu_assign_thr=100).assign(A_, G_).gt_inds.shape == (3000,)


def test_massive_ties_take_the_sweep(cuda_dev):
    """More exact ties than the tie list of pass 1 can hold: the library must fall back to the exact tie sweep."""
    import r3det_b200 as R
    rng = np.random.default_rng(7)
    gt = rand_obb(6, 11, "v1", 30, 200)
    reps = 40000                                              # 6 x 40000 tie candidates > list capacity (65536 + 64 G)
    an = np.concatenate([np.repeat(gt, reps, axis=0), rand_obb(5000, 12, "v1")]).astype(np.float32)
    an = an[rng.permutation(len(an))]
    G_, A_ = torch.from_numpy(gt).to(cuda_dev), torch.from_numpy(an).to(cuda_dev)
    ov = R.pairwise_iou(G_, A_, "v1")
    for pos, minpos in ((0.5, 0.0), (1.5, 0.3)):              # pos 1.5: only low-quality matching assigns positives
        want, _ = assign_wrt_overlaps(ov, pos, 0.4, minpos, True, True)
        out = R.max_iou_assign(G_, A_, pos, 0.4, minpos, True, True, "v1")
        assert torch.equal(out.gt_inds, want), int((out.gt_inds != want).sum())


@pytest.mark.parametrize("v", ["v1", "v3"])
@pytest.mark.parametrize("shared", [True, False])
def test_batched_equals_per_image(cuda_dev, v, shared):
    """One launch sequence for a batch of images == the per-image calls, bit for bit (shared anchors: the anchor-head stage;
    per-image boxes: the refine stage); images without GT and GT counts that are not multiples of the row tile included."""
    import r3det_b200 as R
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda_dev)
    counts = [37, 0, 130, 1, 64]
    gts = [t(rand_obb(c, 50 + i, v, 10, 300)) for i, c in enumerate(counts)]
    A = 5003
    if shared:
        anchors = t(rand_obb(A, 9, v))
    else:
        anchors = t(np.stack([rand_obb(A, 20 + i, v) for i in range(len(counts))]))
    for lq, aa in ((True, True), (True, False), (False, True)):
        got = R.max_iou_assign_batched(gts, anchors, 0.5, 0.4, 0.0, lq, aa, v)
        assert len(got) == len(counts)
        for b, c in enumerate(counts):
            want = R.max_iou_assign(gts[b], anchors if shared else anchors[b], 0.5, 0.4, 0.0, lq, aa, v)
            assert got[b].num_gts == c
            for k in ("gt_inds", "max_overlaps", "argmax_overlaps", "gt_max_overlaps", "gt_argmax_overlaps"):
                assert torch.equal(getattr(got[b], k), getattr(want, k)), (v, shared, lq, aa, b, k)
    # exact ties across a batch (the symmetric set of test_against_matrix_reduction) + the tie-list overflow in one image
    g = t(np.array([[100, 100, 40, 40, 0.0]], np.float32))
    an = t(np.array([[90, 100, 40, 40, 0.0], [110, 100, 40, 40, 0.0], [100, 90, 40, 40, 0.0], [300, 300, 10, 10, 0.0]], np.float32))
    got = R.max_iou_assign_batched([g, g[:0], g], an, 0.9, 0.4, 0.0, True, True, "v1")
    want = R.max_iou_assign(g, an, 0.9, 0.4, 0.0, True, True, "v1")
    assert torch.equal(got[0].gt_inds, want.gt_inds) and torch.equal(got[2].gt_inds, want.gt_inds) and (got[1].gt_inds == 0).all()
    big = rand_obb(3, 11, "v1", 30, 200)
    many = t(np.concatenate([np.repeat(big, 30000, axis=0), rand_obb(2000, 12, "v1")]).astype(np.float32))
    got = R.max_iou_assign_batched([t(big), t(big[:1])], many, 1.5, 0.4, 0.3, True, True, "v1")
    for b, gb in enumerate((t(big), t(big[:1]))):
        assert torch.equal(got[b].gt_inds, R.max_iou_assign(gb, many, 1.5, 0.4, 0.3, True, True, "v1").gt_inds)


def test_batched_randomised(cuda_dev):
    """Random batch shapes (1..9 images, 0..300 GT each, 1..9000 anchors, shared or per-image, every low-quality mode)."""
    import r3det_b200 as R
    rng = np.random.default_rng(21)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(cuda_dev)
    for it in range(14):
        v = ["v1", "v2", "v3"][it % 3]
        B = int(rng.integers(1, 10))
        counts = [int(rng.choice([0, 1, 2, 63, 64, 65, 128, 300])) for _ in range(B)]
        A = int(rng.choice([1, 127, 128, 129, 2000, 9000]))
        shared = bool(rng.integers(2))
        gts = [t(rand_obb(c, int(rng.integers(1 << 30)), v, 10, 300)) if c else torch.zeros((0, 5), device=cuda_dev) for c in counts]
        anchors = t(rand_obb(A, int(rng.integers(1 << 30)), v)) if shared else t(np.stack([rand_obb(A, int(rng.integers(1 << 30)), v) for _ in range(B)]))
        pos, neg, minpos = [(0.5, 0.4, 0.0), (0.6, 0.5, 0.3), (0.3, 0.3, 0.1)][it % 3]
        lq, aa = bool(rng.integers(2)), bool(rng.integers(2))
        got = R.max_iou_assign_batched(gts, anchors, pos, neg, minpos, lq, aa, v)
        for b in range(B):
            want = R.max_iou_assign(gts[b], anchors if shared else anchors[b], pos, neg, minpos, lq, aa, v)
            for k in ("gt_inds", "max_overlaps", "argmax_overlaps", "gt_max_overlaps", "gt_argmax_overlaps"):
                assert torch.equal(getattr(got[b], k), getattr(want, k)), (it, v, B, counts, A, shared, lq, aa, b, k)
