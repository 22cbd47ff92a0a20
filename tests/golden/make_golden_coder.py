"""Generate tests/golden/coder_ref.npz from the REFERENCE'S OWN PYTHON (run in the build container only):

  * r3det/core/bbox/coder/delta_xywha_rbbox_coder.py  (DeltaXYWHAOBBoxCoder.encode / .decode, v1 / v2 / v3)
  * r3det/models/dense_heads/rotate_anchor_head.py    (RAnchorHead._get_bboxes_single incl. multiclass_nms_rotated)
  * r3det/models/dense_heads/rotate_retina_head.py    (RRetinaHead.filter_bboxes)
  * r3det/models/dense_heads/rotate_retina_refine_head.py (RRetinaRefineHead.refine_bboxes)

imported from /root/reference with mmcv / mmdet replaced by import stubs (decorators -> identity, registries -> no-op)
and the native NMS modules replaced by adapters onto the unmodified reference ops compiled in oracle/_ref
(see make_golden.py).  Nothing of the reference's arithmetic is restated here.

    python tests/golden/make_golden_coder.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
AR, rand_obb, AttrDict = mg.AR, mg.rand_obb, mg.AttrDict


def _identity_decorator(*a, **k):
    def deco(f):
        return f
    return deco


class _Registry:
    def register_module(self, *a, **k):
        return lambda c: c


def load_heads_and_coder():
    py = mg.load_reference_python()              # stubs mmcv, r3det, r3det.ops; loads rtransforms + bbox_nms_rotated
    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod("mmcv", jit=_identity_decorator)
    mod("mmcv.cnn", normal_init=None, ConvModule=None, bias_init_with_prob=None)
    mod("mmcv.runner", force_fp32=_identity_decorator)
    mod("mmdet"); mod("mmdet.core", build_assigner=None, build_bbox_coder=None, build_prior_generator=None, build_sampler=None,
                      images_to_levels=None, multi_apply=None, unmap=None)
    mod("mmdet.core.bbox"); mod("mmdet.core.bbox.builder", BBOX_CODERS=_Registry())
    mod("mmdet.core.bbox.coder"); mod("mmdet.core.bbox.coder.base_bbox_coder", BaseBBoxCoder=type("BaseBBoxCoder", (), {}))
    mod("mmdet.models"); mod("mmdet.models.builder", HEADS=_Registry(), build_loss=None)
    mod("mmdet.models.dense_heads")
    mod("mmdet.models.dense_heads.base_dense_head", BaseDenseHead=type("BaseDenseHead", (torch.nn.Module,), {}))
    mod("r3det.core", multiclass_nms_rotated=py["bbox_nms_rotated"].multiclass_nms_rotated,
        obb2hbb=py["rtransforms"].obb2hbb, ranchor_inside_flags=None)

    def load(name, path, package=None):
        sp = importlib.util.spec_from_file_location(name, f"{REF}/{path}")
        m = importlib.util.module_from_spec(sp)
        if package:
            m.__package__ = package
        sys.modules[name] = m
        sp.loader.exec_module(m)
        return m
    coder = load("refpy_coder", "r3det/core/bbox/coder/delta_xywha_rbbox_coder.py")
    pkg = mod("r3det.models"); pkg.__path__ = []
    dh = mod("r3det.models.dense_heads"); dh.__path__ = []
    ah = load("r3det.models.dense_heads.rotate_anchor_head", "r3det/models/dense_heads/rotate_anchor_head.py", "r3det.models.dense_heads")
    rh = load("r3det.models.dense_heads.rotate_retina_head", "r3det/models/dense_heads/rotate_retina_head.py", "r3det.models.dense_heads")
    dh.RRetinaHead = rh.RRetinaHead
    rr = load("r3det.models.dense_heads.rotate_retina_refine_head", "r3det/models/dense_heads/rotate_retina_refine_head.py",
              "r3det.models.dense_heads")
    return coder, ah, rh, rr


def grid_anchors(h, w, stride, num_anchors, rng, v):
    """(H*W*A, 5) anchors in the layout the heads index: location-major, anchor-minor."""
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    ctr = np.stack([xs, ys], -1).reshape(-1, 1, 2).astype(np.float32) * stride + stride / 2
    wh = np.exp(rng.uniform(np.log(2 * stride), np.log(8 * stride), (1, num_anchors, 2))).astype(np.float32)
    ang = rng.uniform(*AR[v], (1, num_anchors, 1)).astype(np.float32) if v != "v1" else np.zeros((1, num_anchors, 1), np.float32)
    a = np.concatenate([np.broadcast_to(ctr, (h * w, num_anchors, 2)), np.broadcast_to(wh, (h * w, num_anchors, 2)),
                        np.broadcast_to(ang, (h * w, num_anchors, 1))], -1)
    return np.ascontiguousarray(a.reshape(-1, 5), np.float32)


def main():
    coder_mod, ah, rh, rr = load_heads_and_coder()
    rng = np.random.default_rng(20260203)
    out = {}
    means, stds = (0.0, 0.0, 0.0, 0.0, 0.0), (1.0, 1.0, 1.0, 1.0, 1.0)
    means2, stds2 = (0.01, -0.02, 0.03, 0.0, 0.05), (0.1, 0.1, 0.2, 0.2, 0.1)
    # ---- coder: encode / decode
    for v in ("v1", "v2", "v3"):
        n = 300
        prop = rand_obb(n, rng, AR[v], 6, 400); gt = prop.copy()
        gt[:, :2] += rng.normal(0, 12, (n, 2)); gt[:, 2:4] *= np.exp(rng.normal(0, 0.4, (n, 2))); gt[:, 4] += rng.normal(0, 0.5, n)
        gt = gt.astype(np.float32)
        deltas = rng.normal(0, 0.6, (n, 5)).astype(np.float32); deltas[::17, 2:4] *= 20         # some hit the wh clip
        deltas3 = rng.normal(0, 0.6, (n, 15)).astype(np.float32)
        out[f"{v}_prop"], out[f"{v}_gt"], out[f"{v}_deltas"], out[f"{v}_deltas3"] = prop, gt, deltas, deltas3
        for tag, (m, s) in (("a", (means, stds)), ("b", (means2, stds2))):
            c = coder_mod.DeltaXYWHAOBBoxCoder(m, s, angle_range=v)
            out[f"{v}_{tag}_encode"] = c.encode(torch.from_numpy(prop), torch.from_numpy(gt)).numpy()
            out[f"{v}_{tag}_decode"] = c.decode(torch.from_numpy(prop), torch.from_numpy(deltas)).numpy()
            out[f"{v}_{tag}_decode3"] = c.decode(torch.from_numpy(prop), torch.from_numpy(deltas3)).numpy()
        c = coder_mod.DeltaXYWHAOBBoxCoder(means2, stds2, angle_range=v)
        out[f"{v}_decode_clamped"] = c.decode(torch.from_numpy(prop), torch.from_numpy(deltas), max_shape=(512, 640, 3)).numpy()
        c = coder_mod.DeltaXYWHAOBBoxCoder(means2, stds2, angle_range=v, add_ctr_clamp=True, ctr_clamp=8)
        out[f"{v}_decode_ctr"] = c.decode(torch.from_numpy(prop), torch.from_numpy(deltas)).numpy()
    out["means_b"], out["stds_b"] = np.array(means2, np.float32), np.array(stds2, np.float32)

    # ---- heads: get_bboxes tail, filter_bboxes, refine_bboxes
    ncls = 15
    sizes = [(16, 20, 8), (8, 10, 16), (4, 5, 32)]                      # (H, W, stride)
    for v in ("v1", "v2", "v3"):
        for A in (1, 3):
            tag = f"{v}_A{A}"
            cls_list, reg_list, anc_list = [], [], []
            for h, w, s in sizes:
                cls = (rng.normal(-3.0, 2.0, (A * ncls, h, w))).astype(np.float32)
                reg = rng.normal(0, 0.4, (A * 5, h, w)).astype(np.float32)
                cls_list.append(cls); reg_list.append(reg); anc_list.append(grid_anchors(h, w, s, A, rng, v))
            coder = coder_mod.DeltaXYWHAOBBoxCoder(means, (1.0, 1.0, 1.0, 1.0, 1.0) if v != "v1" else (0.5, 0.5, 0.5, 0.5, 0.5), angle_range=v)
            fake = types.SimpleNamespace(test_cfg=None, cls_out_channels=ncls, use_sigmoid_cls=True, bbox_coder=coder,
                                         num_anchors=A, anchor_generator=types.SimpleNamespace(
                                             grid_priors=lambda fs, device=None, _a=anc_list: [torch.from_numpy(x) for x in _a]))
            cfg = AttrDict(nms_pre=100, min_bbox_size=0, score_thr=0.05, nms=AttrDict(type=v, iou_thr=0.1), max_per_img=60)
            tl = lambda L: [torch.from_numpy(x) for x in L]
            img_shape = (140, 170, 3); sf = np.array([1.25, 1.5, 1.25, 1.5], np.float32)
            for rescale in (False, True):
                b, s_ = ah.RAnchorHead._get_bboxes_single(fake, tl(cls_list), tl(reg_list), tl(anc_list), img_shape, sf, cfg, rescale, False)
                d, l = ah.RAnchorHead._get_bboxes_single(fake, tl(cls_list), tl(reg_list), tl(anc_list), img_shape, sf, cfg, rescale, True)
                r = int(rescale)
                out[f"{tag}_r{r}_mlvl_bboxes"], out[f"{tag}_r{r}_mlvl_scores"] = b.numpy(), s_.numpy()
                out[f"{tag}_r{r}_dets"], out[f"{tag}_r{r}_labels"] = d.numpy(), l.numpy()
            for i, (c_, r_, a_) in enumerate(zip(cls_list, reg_list, anc_list)):
                out[f"{tag}_cls{i}"], out[f"{tag}_reg{i}"], out[f"{tag}_anc{i}"] = c_, r_, a_
            out[f"{tag}_stds"] = np.array(coder.stds, np.float32)
            # filter_bboxes: batch of 2 images (second image = negated logits / deltas)
            cls_b = [torch.from_numpy(np.stack([x, -x[::-1].copy()])) for x in cls_list]
            reg_b = [torch.from_numpy(np.stack([x, -x])) for x in reg_list]
            fl = rh.RRetinaHead.filter_bboxes(fake, cls_b, reg_b)
            for img in range(2):
                for i in range(len(sizes)):
                    out[f"{tag}_filter_img{img}_lvl{i}"] = fl[img][i].numpy()
            if A == 1:
                rois = [[torch.from_numpy(np.ascontiguousarray(fl[img][i].numpy())) for i in range(len(sizes))] for img in range(2)]
                rf = rr.RRetinaRefineHead.refine_bboxes(fake, cls_b, reg_b, rois)
                for img in range(2):
                    for i in range(len(sizes)):
                        out[f"{tag}_refine_img{img}_lvl{i}"] = rf[img][i].numpy()
    np.savez_compressed(os.path.join(HERE, "coder_ref.npz"), **out)
    print("coder_ref.npz", os.path.getsize(os.path.join(HERE, "coder_ref.npz")), len(out), "arrays")


if __name__ == "__main__":
    main()
