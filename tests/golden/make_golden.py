"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

Sources of truth, none of them restated here:
  * the unmodified reference CPU ops compiled into oracle/_ref (oracle/build_ref.py, oracle/ref.py);
  * the reference's own Python: r3det/core/bbox/rtransforms.py and
    r3det/core/post_processing/bbox_nms_rotated.py + the op wrappers (rnms_wrapper.py,
    nms_rotated_wrapper.py), imported from /root/reference with import stubs (mmcv / mmdet / the
    pybind modules are replaced by thin adapters onto oracle/_ref).
The .npz files are small (< 300 KB in total) and committed; /root/reference is NOT needed to run the tests.

    python tests/golden/make_golden.py
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
from oracle import ref  # noqa: E402

REF = "/root/reference"
AR = {'v1': (-np.pi / 2, 0), 'v2': (-np.pi / 4, 3 * np.pi / 4), 'v3': (-np.pi / 2, np.pi / 2)}


def rand_obb(n, rng, ar, lo=8, hi=512, span=1024):
    cx = rng.uniform(0, span, n); cy = rng.uniform(0, span, n)
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n)); h = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    a = rng.uniform(*ar, n)
    return np.stack([cx, cy, w, h, a], 1).astype(np.float32)


def clustered(K, rng, ar, ncls=15):
    seeds = rand_obb(max(K // 10, 1), rng, ar, 12, 200)
    idx = rng.integers(0, len(seeds), K)
    b = seeds[idx].copy()
    b[:, 0:2] += rng.normal(0, 4, (K, 2)); b[:, 4] += rng.normal(0, 0.05, K)
    b[:, 2:4] *= np.exp(rng.normal(0, 0.1, (K, 2)))
    labels = (idx % ncls).astype(np.int64)
    scores = rng.permutation(np.linspace(0.05, 1, K)).astype(np.float32)
    return b.astype(np.float32), scores, labels


def mixed_boxes(rng, ar, n):
    """random + small + axis-aligned + near-duplicates: the regimes the parity sweep covers."""
    a = rand_obb(n // 2, rng, ar)
    b = rand_obb(n // 8, rng, ar, 0.5, 8, 64)
    c = rand_obb(n // 8, rng, ar, 16, 300); c[:, 4] = 0
    d = a[: n // 8].copy(); d[:, :2] += rng.normal(0, 1e-3, (len(d), 2)).astype(np.float32)
    e = rand_obb(n - len(a) - len(b) - len(c) - len(d), rng, ar, 2, 900)
    return np.concatenate([a, b, c, d, e]).astype(np.float32)


# ---------------------------------------------------------------------------------------- reference python
def load_reference_python():
    """Import the reference's pure-Python layers with the native modules stubbed onto oracle/_ref."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _RnmsExt:
        @staticmethod
        def rnms(dets, thr):
            return torch.from_numpy(ref.v1_nms(dets.numpy(), thr))

    class _NmsRotatedExt:
        @staticmethod
        def nms_rotated(dets, scores, thr):
            return torch.from_numpy(ref.v3_nms(dets.numpy(), scores.numpy(), thr))

    def _ml(dets, scores, labels, thr):
        return torch.from_numpy(ref.v2_nms(dets.numpy(), scores.numpy(), labels.numpy(), thr))

    mod("mmcv"); mod("mmcv.ops", nms_rotated=None, box_iou_rotated=None)
    r3 = mod("r3det"); r3.__path__ = []
    ops = mod("r3det.ops", convex_sort=None); ops.__path__ = []
    for pkg, fname, ext_name, ext in [("rnms", "rnms_wrapper.py", "rnms_ext", _RnmsExt),
                                      ("nms_rotated", "nms_rotated_wrapper.py", "nms_rotated_ext", _NmsRotatedExt)]:
        p = mod(f"r3det.ops.{pkg}"); p.__path__ = []
        setattr(p, ext_name, ext)
        sys.modules[f"r3det.ops.{pkg}.{ext_name}"] = ext
        spec = importlib.util.spec_from_file_location(f"r3det.ops.{pkg}.wrapper", f"{REF}/r3det/ops/{pkg}/{fname}")
        m = importlib.util.module_from_spec(spec); m.__package__ = f"r3det.ops.{pkg}"
        spec.loader.exec_module(m)
        for k in ("rnms", "batched_rnms", "obb_nms", "obb_batched_nms"):
            if hasattr(m, k):
                setattr(ops, k, getattr(m, k))
    ops.ml_nms_rotated = _ml
    out = {}
    for key, path in [("rtransforms", "r3det/core/bbox/rtransforms.py"),
                      ("bbox_nms_rotated", "r3det/core/post_processing/bbox_nms_rotated.py")]:
        spec = importlib.util.spec_from_file_location("refpy_" + key, f"{REF}/{path}")
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        out[key] = m
    return out


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def main():
    rng = np.random.default_rng(20260117)
    # ---- IoU / IoF from the reference binaries
    iou = {}
    for v in ("v1", "v2", "v3"):
        b1 = mixed_boxes(rng, AR[v], 96); b2 = mixed_boxes(rng, AR[v], 160)
        c1, _, _ = clustered(64, rng, AR[v]); b1 = np.concatenate([b1, c1]); b2 = np.concatenate([b2, c1 + np.float32(0.25)])
        iou[f"{v}_b1"], iou[f"{v}_b2"] = b1, b2
        if v == "v1":
            iou["v1_iou"] = ref.v1_iou(b1, b2); iou["v1_iof"] = ref.v1_iou(b1, b2, iof=True)
            iou["v1_aligned"] = ref.v1_iou_aligned(b1[:100], b2[:100])
            iou["v1_aligned_bcast"] = ref.v1_iou_aligned(b1[:1], b2[:50])
        elif v == "v3":
            iou["v3_iou"] = ref.v3_iou(b1, b2, True, through_tensor=True); iou["v3_iof"] = ref.v3_iou(b1, b2, False)
        else:
            iou["v2_iou"] = ref.v2_iou(b1, b2)
    # the survey's known-answer pair
    iou["kat_b1"] = np.array([[50, 50, 40, 20, 0.3]], np.float32); iou["kat_b2"] = np.array([[55, 48, 30, 25, -0.5]], np.float32)
    iou["kat_v3_iou"] = ref.v3_iou(iou["kat_b1"], iou["kat_b2"], True); iou["kat_v3_iof"] = ref.v3_iou(iou["kat_b1"], iou["kat_b2"], False)
    np.savez_compressed(os.path.join(HERE, "iou_ref.npz"), **iou)

    # ---- NMS from the reference binaries (CPU rule: >=)
    nms = {}
    for v in ("v1", "v2", "v3"):
        b, s, l = clustered(700, rng, AR[v])
        nms[f"{v}_boxes"], nms[f"{v}_scores"], nms[f"{v}_labels"] = b, s, l
        if v == "v1":
            nms["v1_keep"] = ref.v1_nms(np.concatenate([b, s[:, None]], 1), 0.1)
        elif v == "v3":
            nms["v3_keep"] = ref.v3_nms(b, s, 0.1)
        else:
            nms["v2_keep"] = ref.v2_nms(b, s, l, 0.1)
    np.savez_compressed(os.path.join(HERE, "nms_ref.npz"), **nms)

    # ---- reference Python: transforms + multiclass_nms_rotated (CPU tensors)
    py = load_reference_python()
    rt = py["rtransforms"]
    tf = {}
    for v in ("v1", "v2", "v3"):
        obb = rand_obb(200, rng, AR[v], 4, 600)
        t = torch.from_numpy(obb)
        poly = rt.obb2poly(t, v)
        tf[f"{v}_obb"] = obb
        tf[f"{v}_obb2poly"] = poly.numpy()
        tf[f"{v}_poly2obb"] = rt.poly2obb(poly, v).numpy()
        # a generic (non-rectangular order) polygon set as well
        rp = (torch.from_numpy(rand_obb(200, rng, AR[v], 4, 600))); rpoly = rt.obb2poly(rp, 'v1').roll(2, 1)
        tf[f"{v}_rpoly"] = rpoly.numpy(); tf[f"{v}_rpoly2obb"] = rt.poly2obb(rpoly, v).numpy()
        tf[f"{v}_obb2hbb"] = rt.obb2hbb(t, v).numpy()
        tf[f"{v}_obb2xyxy"] = rt.obb2xyxy(t, v).numpy()
        hbb = rt.obb2xyxy(t, 'v3')
        tf[f"{v}_hbb"] = hbb.numpy(); tf[f"{v}_hbb2obb"] = rt.hbb2obb(hbb, v).numpy()
    np.savez_compressed(os.path.join(HERE, "transforms_ref.npz"), **tf)

    mc = {}
    fn = py["bbox_nms_rotated"].multiclass_nms_rotated
    for v in ("v1", "v2", "v3"):
        n, ncls = 400, 15
        b, _, _ = clustered(n, rng, AR[v])
        sc = (rng.uniform(0, 1, (n, ncls + 1)) ** 6).astype(np.float32)       # sparse high scores
        flat = sc[:, :-1].reshape(-1); flat[:] = rng.permutation(np.linspace(0.0, 1.0, flat.size)).astype(np.float32) ** 6
        sc[:, :-1] = flat.reshape(n, ncls)
        for max_num in (50, 2000):
            dets, labels = fn(torch.from_numpy(b), torch.from_numpy(sc), 0.05, AttrDict(type=v, iou_thr=0.1), max_num)
            mc[f"{v}_{max_num}_dets"], mc[f"{v}_{max_num}_labels"] = dets.numpy(), labels.numpy()
        mc[f"{v}_boxes"], mc[f"{v}_scores"] = b, sc
    # empty case
    dets, labels = fn(torch.from_numpy(b), torch.zeros(n, ncls + 1), 0.05, AttrDict(type='v1', iou_thr=0.1), 2000)
    mc["empty_dets_shape"] = np.array(dets.shape); mc["empty_labels_shape"] = np.array(labels.shape)
    np.savez_compressed(os.path.join(HERE, "multiclass_ref.npz"), **mc)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
