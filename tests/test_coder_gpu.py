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


def _torch_decode(rois, deltas, means, stds, version, max_shape=None, wh_ratio_clip=16 / 1000, add_ctr_clamp=False, ctr_clamp=32):
    """The reference's torch expressions (delta_xywha_rbbox_coder.py:172-211, 283-311, 391-423), kept differentiable."""
    import math
    g = deltas.size(1) // 5
    den = deltas * deltas.new_tensor(stds).repeat(g)[None] + deltas.new_tensor(means).repeat(g)[None]
    dx, dy, dw, dh, da = (den[:, i::5] for i in range(5))
    px, py, pw, ph, pa = (rois[:, i, None].expand_as(dx) for i in range(5))
    mr = abs(math.log(wh_ratio_clip))
    if version == 'v1':
        sx, sy = pw * dx, ph * dy
        if add_ctr_clamp:
            sx, sy = sx.clamp(-ctr_clamp, ctr_clamp), sy.clamp(-ctr_clamp, ctr_clamp)
            dw, dh = dw.clamp(max=mr), dh.clamp(max=mr)
        else:
            dw, dh = dw.clamp(-mr, mr), dh.clamp(-mr, mr)
        gx, gy, gw, gh, ga = px + sx, py + sy, pw * dw.exp(), ph * dh.exp(), pa + da
        if max_shape is not None:
            gx, gy = gx.clamp(0, max_shape[1] - 1), gy.clamp(0, max_shape[0] - 1)
        return torch.stack([gx, gy, gw, gh, ga], -1).view(deltas.size())
    dw, dh = dw.clamp(-mr, mr), dh.clamp(-mr, mr)
    ang = pa if version == 'v2' else -pa
    gx = dx * pw * torch.cos(ang) - dy * ph * torch.sin(ang) + px
    gy = dx * pw * torch.sin(ang) + dy * ph * torch.cos(ang) + py
    gw, gh = pw * dw.exp(), ph * dh.exp()
    if version == 'v2':
        gt = (da * math.pi + pa + math.pi / 4) % math.pi - math.pi / 4
        return torch.stack([gx, gy, gw, gh, gt], -1).view_as(deltas)
    gt = da + pa
    wr, hr = torch.where(gw > gh, gw, gh), torch.where(gw > gh, gh, gw)
    tr = (torch.where(gw > gh, gt, gt + math.pi / 2) + math.pi / 2) % math.pi - math.pi / 2
    return torch.stack([gx, gy, wr, hr, tr], -1).view_as(deltas)


@pytest.mark.parametrize("version", ["v1", "v2", "v3"])
def test_decode_is_differentiable(cuda_dev, version):
    """rotate_anchor_head.py:418-420 back-propagates through bbox_coder.decode (reg_decoded_bbox=True); the gradient
    with respect to the deltas must be the reference's (torch autograd over its own expressions), not silently zero."""
    import r3det_b200 as R
    rng = np.random.default_rng(3)
    rois = torch.from_numpy(rand_obb(400, 5, version)).to(cuda_dev)
    means, stds = (0.01, -0.02, 0.03, 0.0, 0.05), (0.1, 0.2, 0.3, 0.4, 0.5)
    cases = [dict()]
    if version == 'v1':
        cases += [dict(max_shape=(600, 700)), dict(add_ctr_clamp=True, ctr_clamp=8)]
    for groups in (1, 3):
        raw = rng.normal(0, 3.0, (400, 5 * groups)).astype(np.float32)          # wide: the size clamps are active for some rows
        for kw in cases:
            coder = R.DeltaXYWHAOBBoxCoder(means, stds, angle_range=version, add_ctr_clamp=kw.get('add_ctr_clamp', False),
                                           ctr_clamp=kw.get('ctr_clamp', 32))
            d1 = torch.from_numpy(raw).to(cuda_dev).requires_grad_(True)
            d2 = torch.from_numpy(raw).to(cuda_dev).requires_grad_(True)
            if groups == 1:
                out = coder.decode(rois, d1, max_shape=kw.get('max_shape'))
            else:
                out = getattr(R, f'delta2bbox_{version}')(rois, d1, means, stds, **({'max_shape': kw.get('max_shape')} if version == 'v1' else {}),
                                                           **({k: v for k, v in kw.items() if k != 'max_shape'} if version == 'v1' else {}))
            ref = _torch_decode(rois, d2, means, stds, version, **kw)
            assert out.requires_grad and out.grad_fn is not None
            wgt = torch.from_numpy(rng.normal(0, 1, out.shape).astype(np.float32)).to(cuda_dev)
            (out * wgt).sum().backward()
            (ref * wgt).sum().backward()
            scale = float(d2.grad.abs().max())
            assert float((d1.grad - d2.grad).abs().max()) <= 1e-4 * max(scale, 1.0), (version, groups, kw)
    # boxes being decoded (anchors) do not get a gradient: asking for one raises instead of returning zero
    with pytest.raises(RuntimeError):
        R.DeltaXYWHAOBBoxCoder(angle_range=version).decode(rois.clone().requires_grad_(True), torch.zeros((400, 5), device=cuda_dev))
    # without grad mode the plain kernel path is taken
    with torch.no_grad():
        o = R.DeltaXYWHAOBBoxCoder(angle_range=version).decode(rois, torch.zeros((400, 5), device=cuda_dev, requires_grad=True))
    assert not o.requires_grad


def test_non_differentiable_ops_raise_on_grad_inputs(cuda_dev):
    import r3det_b200 as R
    b = torch.from_numpy(rand_obb(16, 1, "v1")).to(cuda_dev)
    with pytest.raises(RuntimeError):
        R.obb2poly(b.clone().requires_grad_(True), "v1")
    with pytest.raises(RuntimeError):
        R.DeltaXYWHAOBBoxCoder(angle_range="v1").encode(b.clone().requires_grad_(True), b)
    with pytest.raises(RuntimeError):
        R.obb_overlaps(b.clone().requires_grad_(True), b, is_aligned=True)
    with torch.no_grad():
        R.obb2poly(b.clone().requires_grad_(True), "v1")
    # aligned v3 overlaps apply no too-small mask (the reference's torch aligned path has none)
    tiny = b.clone(); tiny[:, 2] = 5e-4
    assert float(R.obb_overlaps(tiny, tiny, is_aligned=True).min()) > 0.5
    assert float(R.obb_overlaps(tiny, tiny).max()) == 0.0
