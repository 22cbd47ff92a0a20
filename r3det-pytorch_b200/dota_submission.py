"""Offline DOTA Task-1 submission: merge patch detections into full images with rotated / polygon NMS on the GPU and
write the `Task1_<class>.txt` files — mirror of DOTADataset.merge_det / _merge_func / _results2submission
(r3det/datasets/dota1.py:209-292, 632-667), the on-disk side of the NMS path.

The reference runs one NMS call per (image, class) from a Python loop; here ALL classes of an image go through one
launch sequence (labels are segments of the same call), the polygon variant included."""
import os
import re
import zipfile
from collections import defaultdict

import numpy as np
import torch

from ._nms_core import nms_device
from .nms_rotated import poly_nms_device
from .rtransforms import obb2poly_np

_PATCH = re.compile(r'__\d+___\d+')


def patch_origin(img_id):
    """'P0006__1__0___824' -> ('P0006', 0, 824)  (dota1.py:218-224)."""
    x, y = (int(v) for v in re.findall(r'\d+', _PATCH.findall(img_id)[0])[:2])
    return img_id.split('__')[0], x, y


def collect_patches(results, img_ids):
    """results[i][c]: (n, 6) detections of class c in patch i -> {image: (N, 7) rows [label, x, y, w, h, a, score]} with the
    patch offset added (dota1.py:215-236)."""
    collector = defaultdict(list)
    for result, img_id in zip(results, img_ids):
        name, x, y = patch_origin(img_id)
        rows = []
        for c, dets in enumerate(result):
            dets = np.asarray(dets)
            boxes = dets[:, :-1].copy()
            boxes[..., :2] = boxes[..., :2] + np.array([x, y], dtype=np.float32)
            rows.append(np.concatenate([np.zeros((boxes.shape[0], 1)) + c, boxes, dets[:, [-1]]], axis=1))
        collector[name].append(np.concatenate(rows, axis=0))
    return {k: np.concatenate(v, axis=0) for k, v in collector.items()}


def merge_image(label_dets, num_classes, iou_thr=0.1, version='v1', merge_nms='obb', device=None):
    """_merge_func (dota1.py:632-667) for one image: (N, 7) rows -> list over classes of kept (k, 6) detections.
    merge_nms 'poly' -> polygon NMS of obb2poly_np(dets, version) (rule >); otherwise rnms (v1: rule >=, kept rows in
    ascending index) or obb_nms (v2 / v3: rule >=, descending score) exactly as the reference's numpy calls resolve.
    PRECISION: the reference passes these float64 rows to its CPU ops (double templates) and rotates obb2poly_np in float64;
    here the rows are cast to float32 for the device kernels.  On full-image coordinates (~1e4 px, FP32 ulp 1e-3 px) the IoU of a
    pair can move by ~1e-4, so a pair whose float64 IoU lies that close to iou_thr may be decided differently; the returned rows
    themselves are the caller's float64 rows, untouched (tests/test_nms_gpu.py::test_dota_merge_float64_full_image_coordinates)."""
    device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    label_dets = np.asarray(label_dets)
    labels, dets = label_dets[:, 0], label_dets[:, 1:]
    out = []
    if merge_nms == 'poly':
        if len(dets) == 0:
            return [dets[labels == c] for c in range(num_classes)]
        polys = torch.from_numpy(np.ascontiguousarray(obb2poly_np(dets, version), np.float32)).to(device)
        lab = torch.from_numpy(labels.astype(np.int64)).to(device)
        keep, num = poly_nms_device(polys[:, :8], polys[:, 8], iou_thr, labels=lab)        # every class in one call
        keep = keep[:int(num.item())].cpu().numpy()                                        # descending score
        kept_labels = labels[keep]
        return [dets[keep[kept_labels == c]] for c in range(num_classes)]
    if len(dets) == 0:
        return [dets[labels == c] for c in range(num_classes)]
    d = torch.from_numpy(np.ascontiguousarray(dets, np.float32)).to(device)
    lab = torch.from_numpy(labels.astype(np.int64)).to(device)
    v1 = version == 'v1'
    keep, num = nms_device(d[:, :5], d[:, 5], iou_thr, 'v1' if v1 else 'v3', labels=lab, inclusive=True,
                           order_index=v1, drop_small=not v1)
    keep = keep[:int(num.item())].cpu().numpy()
    kept_labels = labels[keep]
    return [dets[keep[kept_labels == c]] for c in range(num_classes)]


def merge_det(results, img_ids, classes, iou_thr=0.1, version='v1', merge_nms='obb', device=None):
    """DOTADataset.merge_det: (image ids, per-image list of per-class arrays)."""
    merged = {k: merge_image(v, len(classes), iou_thr, version, merge_nms, device) for k, v in collect_patches(results, img_ids).items()}
    return list(merged.keys()), list(merged.values())


def write_task1(out_folder, id_list, dets_list, classes, version='v1'):
    """_results2submission (dota1.py:250-292): one `Task1_<class>.txt` per class with lines
    `<image> <score> x0 y0 x1 y1 x2 y2 x3 y3` (coordinates %.2f), zipped next to them.  Returns the file list."""
    if os.path.exists(out_folder):
        raise ValueError(f'The out_folder should be a non-exist path, but {out_folder} is existing')
    os.makedirs(out_folder)
    files = [os.path.join(out_folder, 'Task1_' + cls + '.txt') for cls in classes]
    handles = [open(f, 'w') for f in files]
    try:
        for img_id, per_cls in zip(id_list, dets_list):
            for fh, dets in zip(handles, per_cls):
                if len(dets) == 0:
                    continue
                for row in obb2poly_np(dets, version):
                    fh.write(' '.join([img_id, str(row[-1])] + [f'{p:.2f}' for p in row[:-1]]) + '\n')
    finally:
        for fh in handles:
            fh.close()
    name = os.path.split(out_folder)[-1]
    with zipfile.ZipFile(os.path.join(out_folder, name + '.zip'), 'w', zipfile.ZIP_DEFLATED) as z:
        for f in files:
            z.write(f, os.path.split(f)[-1])
    return files
