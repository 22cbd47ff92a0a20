"""GPU parity: rotated box coder, the get_bboxes tail (select + decode + multiclass NMS), filter_bboxes and
refine_bboxes against (a) vectors produced by the reference's own Python (tests/golden/coder_ref.npz) and (b) the numpy
oracle (oracle/coder_np.py) at detection sizes.  Tolerance 2e-5 relative (exp / log / sin / cos differ by an ulp or two
between CUDA libm, torch-CPU and numpy); integer outputs (labels, row selection) exact."""
import numpy as np
import pytest
import torch

from oracle import coder_np as cn
from tests.util import golden, rand_obb

pytestmark = pytest.mark.gpu
ZERO, ONE = (0.0,) * 5, (1.0,) * 5


def close(a, b, tol=2e-5):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    a = a.astype(np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= tol, err.max()


class Cfg(dict):
    __getattr__ = dict.__getitem__


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_coder_golden(cuda_dev, v):
    import r3det_b200 as R
    g = golden("coder_ref.npz")
    t = lambda k: torch.from_numpy(g[k]).to(cuda_dev)
    mb, sb = tuple(g["means_b"].tolist()), tuple(g["stds_b"].tolist())
    for tag, (m, s) in (("a", (ZERO, ONE)), ("b", (mb, sb))):
        c = R.DeltaXYWHAOBBoxCoder(m, s, angle_range=v)
        close(c.encode(t(f"{v}_prop"), t(f"{v}_gt")), g[f"{v}_{tag}_encode"], 5e-5)
        close(c.decode(t(f"{v}_prop"), t(f"{v}_deltas")), g[f"{v}_{tag}_decode"])
        close(c.decode(t(f"{v}_prop"), t(f"{v}_deltas3")), g[f"{v}_{tag}_decode3"])
    c = R.DeltaXYWHAOBBoxCoder(mb, sb, angle_range=v)
    close(c.decode(t(f"{v}_prop"), t(f"{v}_deltas"), max_shape=(512, 640, 3)), g[f"{v}_decode_clamped"])
    c = R.DeltaXYWHAOBBoxCoder(mb, sb, angle_range=v, add_ctr_clamp=True, ctr_clamp=8)
    close(c.decode(t(f"{v}_prop"), t(f"{v}_deltas")), g[f"{v}_decode_ctr"])
    fn = {"v1": (R.bbox2delta_v1, R.delta2bbox_v1), "v2": (R.bbox2delta_v2, R.delta2bbox_v2), "v3": (R.bbox2delta_v3, R.delta2bbox_v3)}[v]
    close(fn[0](t(f"{v}_prop"), t(f"{v}_gt"), mb, sb), g[f"{v}_b_encode"], 5e-5)
    close(fn[1](t(f"{v}_prop"), t(f"{v}_deltas"), mb, sb), g[f"{v}_b_decode"])


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_coder_vs_oracle_large_and_roundtrip(cuda_dev, v):
    import r3det_b200 as R
    rng = np.random.default_rng(5)
    n = 200000
    prop = rand_obb(n, 1, v, 4, 600); gt = prop.copy()
    gt[:, :2] += rng.normal(0, 20, (n, 2)); gt[:, 2:4] *= np.exp(rng.normal(0, 0.5, (n, 2))); gt[:, 4] += rng.normal(0, 0.6, n)
    gt = gt.astype(np.float32)
    m, s = (0.0, 0.01, 0.0, -0.01, 0.0), (0.1, 0.1, 0.2, 0.2, 0.1)
    c = R.DeltaXYWHAOBBoxCoder(m, s, angle_range=v)
    P, G = torch.from_numpy(prop).to(cuda_dev), torch.from_numpy(gt).to(cuda_dev)
    enc = c.encode(P, G)
    want = cn.bbox2delta(prop, gt, m, s, v)
    # the angle wrap is discontinuous: a last-ulp difference in (ga - pa + off) % pi may flip a few rows by pi
    bad = np.abs(enc.cpu().numpy() - want).max(1) > 1e-3
    assert bad.mean() < 1e-4
    close(enc.cpu().numpy()[~bad], want[~bad], 1e-4)
    dec = c.decode(P, enc)
    wantd = cn.delta2bbox(prop, enc.cpu().numpy(), m, s, v)
    badd = np.abs(dec.cpu().numpy() - wantd).max(1) > 1e-3
    assert badd.mean() < 1e-4
    close(dec.cpu().numpy()[~badd], wantd[~badd], 5e-5)
    if v == "v1":                                            # encode -> decode is the identity for v1 (no re-ordering of w / h)
        close(dec, gt, 2e-4)
    assert c.decode(P[:0], enc[:0]).shape == (0, 5) and c.encode(P[:0], G[:0]).shape == (0, 5)
    with pytest.raises(RuntimeError):
        c.decode(P.cpu(), enc.cpu())


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
@pytest.mark.parametrize("A", [1, 3])
def test_heads_golden(cuda_dev, v, A):
    import r3det_b200 as R
    g = golden("coder_ref.npz")
    tag = f"{v}_A{A}"
    coder = R.DeltaXYWHAOBBoxCoder(ZERO, tuple(g[f"{tag}_stds"].tolist()), angle_range=v)
    cls = [g[f"{tag}_cls{i}"] for i in range(3)]; reg = [g[f"{tag}_reg{i}"] for i in range(3)]; anc = [g[f"{tag}_anc{i}"] for i in range(3)]
    cls2 = [np.stack([x, -x[::-1]]) for x in cls]; reg2 = [np.stack([x, -x]) for x in reg]
    T = lambda L_: [torch.from_numpy(np.ascontiguousarray(x)).to(cuda_dev) for x in L_]
    cfg = Cfg(nms_pre=100, min_bbox_size=0, score_thr=0.05, nms=Cfg(type=v, iou_thr=0.1), max_per_img=60)
    metas = [dict(img_shape=(140, 170, 3), scale_factor=np.array([1.25, 1.5, 1.25, 1.5], np.float32))] * 2
    for r in (0, 1):
        raw = R.get_bboxes(T(cls2), T(reg2), T(anc), metas, cfg, coder, rescale=bool(r), with_nms=False)
        close(raw[0][0], g[f"{tag}_r{r}_mlvl_bboxes"]); close(raw[0][1], g[f"{tag}_r{r}_mlvl_scores"], 1e-6)
        dets = R.get_bboxes(T(cls2), T(reg2), T(anc), metas, cfg, coder, rescale=bool(r), with_nms=True)
        d, l = dets[0]
        assert np.array_equal(l.cpu().numpy(), g[f"{tag}_r{r}_labels"])
        close(d, g[f"{tag}_r{r}_dets"])
    fl = R.filter_bboxes(T(cls2), T(reg2), T(anc), coder)
    for img in range(2):
        for i in range(3):
            close(fl[img][i], g[f"{tag}_filter_img{img}_lvl{i}"])
    if A == 1:
        rois = [[torch.from_numpy(g[f"{tag}_filter_img{img}_lvl{i}"]).to(cuda_dev) for i in range(3)] for img in range(2)]
        rf = R.refine_bboxes(T(cls2), T(reg2), rois, coder)
        for img in range(2):
            for i in range(3):
                close(rf[img][i], g[f"{tag}_refine_img{img}_lvl{i}"])
        rb = R.refine_bboxes(T(cls2), T(reg2), [torch.stack([rois[0][i], rois[1][i]]) for i in range(3)], coder, as_batch=True)
        assert all(torch.equal(rb[i][img], rf[img][i]) for i in range(3) for img in range(2))


def _level_inputs(rng, B, A, C, h, w, stride, v):
    """logits whose per-row maxima are well separated (top-k order is then unambiguous across libms)."""
    n = h * w * A
    cls = rng.normal(-8.0, 1.0, (B, n, C)).astype(np.float32)
    for b in range(B):
        top = rng.permutation(np.linspace(-6.0, 4.0, n)).astype(np.float32)
        cls[b, np.arange(n), rng.integers(0, C, n)] = top
    cls = cls.reshape(B, h, w, A * C).transpose(0, 3, 1, 2)            # row (hw*A + a), class c -> channel a*C + c
    reg = rng.normal(0, 0.3, (B, A * 5, h, w)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    ctr = (np.stack([xs, ys], -1).reshape(-1, 1, 2) * stride + stride / 2).astype(np.float32)
    wh = np.exp(rng.uniform(np.log(2 * stride), np.log(8 * stride), (1, A, 2))).astype(np.float32)
    ang = rng.uniform(-1.5, 0.0, (1, A, 1)).astype(np.float32)
    anc = np.concatenate([np.broadcast_to(ctr, (h * w, A, 2)), np.broadcast_to(wh, (h * w, A, 2)), np.broadcast_to(ang, (h * w, A, 1))], -1)
    return np.ascontiguousarray(cls), reg, np.ascontiguousarray(anc.reshape(-1, 5), np.float32)


@pytest.mark.parametrize("v,A", [("v1", 9), ("v3", 1)])
def test_select_decode_vs_oracle_detection_size(cuda_dev, v, A):
    """5 FPN levels of a 512 x 512 patch (A = 9: 49k rows/image; A = 1 with per-image rois: the refine stage)."""
    import r3det_b200 as R
    rng = np.random.default_rng(11)
    B, C = 3, 15
    lv = [_level_inputs(rng, B, A, C, s, s, 512 // s, v) for s in (64, 32, 16, 8, 4)]
    cls = [x[0] for x in lv]; reg = [x[1] for x in lv]; anc = [x[2] for x in lv]
    if A == 1:                                                          # per-image anchors
        anc = [np.stack([a, a + np.float32(1.5), a * np.float32(1.01)]) for a in anc]
    coder = R.DeltaXYWHAOBBoxCoder(ZERO, ONE, angle_range=v)
    T = lambda L_: [torch.from_numpy(x).to(cuda_dev) for x in L_]
    shapes = [(512, 512, 3), (400, 500, 3), (512, 300, 3)]
    sfs = [np.array([1, 1, 1, 1], np.float32), np.array([0.5, 0.75, 0.5, 0.75], np.float32), np.array([2, 2, 2, 2], np.float32)]
    boxes, scores = R.select_decode(T(cls), T(reg), T(anc), coder, 2000, shapes, sfs)
    cd = dict(means=ZERO, stds=ONE, variant=v)
    for b in range(B):
        wb, wsc = cn.select_decode([c[b] for c in cls], [r[b] for r in reg], [a[b] if a.ndim == 3 else a for a in anc],
                                   shapes[b], sfs[b], 2000, C, cd, rescale=True)
        close(scores[b], wsc, 1e-6)
        close(boxes[b], wb)
    # nms_pre <= 0 keeps every row in its original order
    allb, alls = R.select_decode(T(cls), T(reg), T(anc), coder, -1, None, None)
    wb, wsc = cn.select_decode([c[0] for c in cls], [r[0] for r in reg], [a[0] if a.ndim == 3 else a for a in anc], None, None, -1, C, cd)
    close(alls[0], wsc, 1e-6); close(allb[0], wb)


def test_filter_refine_vs_oracle(cuda_dev):
    import r3det_b200 as R
    rng = np.random.default_rng(3)
    B, A, C = 2, 9, 15
    for v in ("v1", "v2"):
        cls, reg, anc = _level_inputs(rng, B, A, C, 40, 56, 8, v)
        coder = R.DeltaXYWHAOBBoxCoder(ZERO, (0.5,) * 5, angle_range=v)
        cd = dict(means=ZERO, stds=(0.5,) * 5, variant=v)
        t = lambda x: torch.from_numpy(x).to(cuda_dev)
        fl = R.filter_bboxes([t(cls)], [t(reg)], [t(anc)], coder, as_batch=True)[0]
        for b in range(B):
            close(fl[b], cn.filter_bboxes(cls[b], reg[b], anc, A, C, cd))
        reg1 = reg[:, :5]
        rf = R.refine_bboxes([t(cls)], [t(np.ascontiguousarray(reg1))], [fl], coder, as_batch=True)[0]
        for b in range(B):
            close(rf[b], cn.refine_bboxes(reg1[b], fl[b].cpu().numpy(), cd))


def test_select_decode_many_images_and_ties(cuda_dev):
    """More images than one C call takes (64): the wrapper chunks; identical scores everywhere (a freshly initialised
    head): the top-k falls back to the lowest row indices, as a stable sort of the reference's keys would."""
    import r3det_b200 as R
    rng = np.random.default_rng(5)
    B, A, C = 70, 2, 4
    cls, reg, anc = _level_inputs(rng, B, A, C, 6, 5, 8, "v1")
    coder = R.DeltaXYWHAOBBoxCoder(ZERO, ONE, angle_range="v1")
    t = lambda x: torch.from_numpy(x).to(cuda_dev)
    boxes, scores = R.select_decode([t(cls)], [t(reg)], [t(anc)], coder, 20, None, None)
    assert boxes.shape == (B, 20, 5) and scores.shape == (B, 20, C + 1)
    cd = dict(means=ZERO, stds=ONE, variant="v1")
    for b in (0, 63, 64, 69):
        wb, ws = cn.select_decode([cls[b]], [reg[b]], [anc], None, None, 20, C, cd)
        close(scores[b], ws, 1e-6); close(boxes[b], wb)
    flat = np.zeros_like(cls)                                              # all logits equal -> all scores tie
    bt, st = R.select_decode([t(flat[:2])], [t(reg[:2])], [t(anc)], coder, 20, None, None)
    wb, _ = cn.select_decode([flat[0]], [reg[0]], [anc], None, None, 20, C, cd)      # stable argsort: rows 0..19
    close(bt[0], wb)
    assert torch.all(st[..., :C] == 0.5) and torch.all(st[..., C] == 0)
    # a batch whose scores never pass score_thr: every image comes back empty
    metas = [dict(img_shape=(48, 40, 3), scale_factor=np.ones(4, np.float32))] * 2
    out = R.get_bboxes([t(flat[:2] - 20)], [t(reg[:2])], [t(anc)], metas, dict(nms_pre=20, score_thr=0.05, nms=dict(type="v1", iou_thr=0.1),
                                                                                  max_per_img=10), coder)
    assert len(out) == 2 and all(d.shape == (0, 6) and l.shape == (0,) for d, l in out)
