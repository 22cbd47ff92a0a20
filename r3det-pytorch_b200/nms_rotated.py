"""v3 rotated NMS — host-side mirror of r3det/ops/nms_rotated/nms_rotated_wrapper.py
(obb_nms :23-54, obb_batched_nms :79-98).  Keep lists are in descending-score order
(nms_rotated_cuda.cu:131-133); boxes with min(w,h) < 1e-3 take no part (:40-46).  CUDA inputs use the
reference GPU rule (IoU > thr), numpy / CPU inputs the reference CPU rule (IoU >= thr,
nms_rotated_cpu.cpp:55); all compute is on the GPU."""
import torch

from ._nms_core import nms_device, to_cuda_input


def obb2hbb(obboxes):
    """[x_ctr, y_ctr, w, h, angle] -> [x_lt, y_lt, x_rb, y_rb] (nms_rotated_wrapper.py:7-20; the same expression as
    rtransforms.obb2xyxy_v3, which is the kernel that runs here)."""
    from .rtransforms import obb2xyxy
    if obboxes.is_cuda:
        return obb2xyxy(obboxes.float(), 'v3').to(obboxes.dtype)
    dev = torch.device('cuda', torch.cuda.current_device())
    return obb2xyxy(obboxes.float().to(dev), 'v3').to(obboxes.dtype).cpu()


def obb_nms(dets, iou_thr, device_id=None):
    """Compute the NMS of oriented bboxes.  dets: (K, 6) <x, y, w, h, a, score>."""
    dets_th, is_numpy, was_host = to_cuda_input(dets, device_id, "dets")
    if dets_th.numel() == 0:
        inds = dets_th.new_zeros(0, dtype=torch.int64)
    else:
        d = dets_th.float()
        keep, num = nms_device(d[:, :5], d[:, 5], iou_thr, "v3", inclusive=was_host, drop_small=True)
        inds = keep[:int(num.item())]
    if is_numpy:
        inds = inds.cpu().numpy()
    elif was_host:
        inds = inds.cpu()
    return dets[inds, :], inds


def poly_nms_device(polys, scores, iou_thr, labels=None):
    """Polygon NMS on CUDA tensors: polys (K, >=8), scores (K,) -> (keep (K,) int64, num_keep 0-dim int64), both on the
    device; no host synchronisation.  labels (K,) int64: class-wise NMS in one call."""
    import ctypes as C
    from . import _lib as L
    L.require_cuda(polys, scores)
    p, stride = L.as_f32_rows(polys, 8)
    s = scores.float().contiguous()
    K = p.size(0)
    keep = torch.empty((K,), dtype=torch.int64, device=p.device)
    num = torch.zeros((), dtype=torch.int64, device=p.device)
    if K == 0:
        return keep, num
    if labels is not None:
        L.require_cuda(labels)
        labels = labels.to(torch.int64).contiguous()
    lib = L.lib()
    nbytes = C.c_size_t(0)
    L.check(lib.r3g_poly_nms_workspace_bytes(K, C.byref(nbytes)))
    ws = L.workspace(nbytes.value, p.device)
    with L.device_guard(p.device):
        L.check(lib.r3g_poly_nms_f32(L.ptr(p), stride, L.ptr(s), L.ptr(labels), K, float(iou_thr), L.ptr(keep), C.c_void_p(num.data_ptr()),
                                     L.ptr(ws), ws.numel(), L.stream_ptr(p.device)))
    return keep, num


def poly_nms(dets, iou_thr, device_id=None):
    """Compute the NMS of polygons (nms_rotated_wrapper.py:57-76).  dets: (K, 9) [x0, y0, ..., x3, y3, score], a CUDA
    tensor or a numpy array with `device_id`; as in the reference there is no CPU implementation."""
    import numpy as np
    if isinstance(dets, torch.Tensor):
        is_numpy = False
        dets_th = dets
    elif isinstance(dets, np.ndarray):
        is_numpy = True
        device = 'cpu' if device_id is None else f'cuda:{device_id}'
        dets_th = torch.from_numpy(dets).to(device)
    else:
        raise TypeError('dets must be eithr a Tensor or numpy array, '
                        f'but got {type(dets)}')
    if dets_th.device == torch.device('cpu'):
        raise NotImplementedError
    d = dets_th.float()
    keep, num = poly_nms_device(d[:, :8], d[:, 8], iou_thr)
    inds = keep[:int(num.item())]
    if is_numpy:
        inds = inds.cpu().numpy()
    return dets[inds, :], inds


def obb_batched_nms(bboxes, scores, inds, nms_thr, class_agnostic=False):
    """Compute the NMS of oriented bboxes in batches (per class id `inds`).
    (N, 4) horizontal boxes: the reference adds the offsets and then calls obb_nms on an (N, 5) tensor, whose
    `dets_th[:, 5]` raises IndexError (nms_rotated_wrapper.py:47, 92-97) — there is no working behaviour to mirror."""
    if bboxes.size(-1) != 5:
        raise NotImplementedError("obb_batched_nms: only (N, 5) oriented boxes are supported (the reference raises IndexError on (N, 4))")
    if class_agnostic or bboxes.shape[0] == 0:
        dets, keep = obb_nms(torch.cat([bboxes, scores[:, None]], -1), nms_thr)
        return torch.cat([bboxes[keep], dets[:, -1:]], -1), keep
    b, _, was_host = to_cuda_input(bboxes, None, "bboxes")
    s = scores.to(b.device)
    lab = inds.to(b.device)
    hbboxes = obb2hbb(b)
    scale = (hbboxes.max() - hbboxes.min()) + 1              # nms_rotated_wrapper.py:85-86
    keep, num = nms_device(b, s, nms_thr, "v3", labels=lab, class_offset=scale, inclusive=was_host, drop_small=True)
    keep = keep[:int(num.item())]
    if was_host:
        keep = keep.cpu()
    bboxes = bboxes[keep]
    scores = scores[keep]
    return torch.cat([bboxes, scores[:, None]], -1), keep
