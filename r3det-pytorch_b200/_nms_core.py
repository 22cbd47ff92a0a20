"""Shared launcher for the rotated-NMS entry point r3g_nms_f32 (include/r3geo.h)."""
import ctypes as C

import torch

from . import _lib as L


def nms_device(boxes, scores, thr, variant, labels=None, class_offset=None, inclusive=False,
               order_index=False, drop_small=False, strict=True, batch_ids=None, n_batches=1, sort_path=False, stats=None, label_bits=0, count=None):
    """Rotated NMS on CUDA tensors.

    boxes (K, >=5) f32, scores (K,) f32, labels (K,) int64 or None, class_offset: 0-dim CUDA f32 tensor or None.
    Returns (keep, num_keep): keep is a (K,) int64 CUDA tensor whose first num_keep (0-dim int64 CUDA tensor)
    entries are the kept original indices; no host synchronisation happens here.

    Multi-image batches (one launch sequence for many images): batch_ids (K,) int64 in [0, n_batches) with the
    candidates concatenated image by image, class_offset a (n_batches,) tensor of per-image scales; num_keep is then a
    (n_batches,) int64 tensor and keep is grouped by image (index order or per-image score order).
    In a batch, labels must lie in [0, 65536) (the segment key packs image and label into 32 bits) and candidates whose
    batch id is outside [0, n_batches) take no part.

    count: optional 0-dim int64 CUDA tensor — only the first `count` candidates are real, the rest is padding written by
    r3g_mc_candidates_batched_f32 (score -inf, image id 65535); needs batch_ids.

    label_bits: optional promise that every label is < 2**label_bits (a caller that knows its class count saves radix-sort
    passes above 16384 candidates); 0 = unknown.

    stats: optional dict; receives `counters` — a (64,) int64 CUDA tensor view of the library's work counters for this
    call (read it after synchronising): [0] stage-1 pair tests, [1] separating-axis tests, [2] area evaluations,
    [3] reference restatements, [4] rounds, [5] mask items, [6] apply items, [8] n, [9..] phase time stamps in ns.
    """
    L.require_cuda(boxes, scores)
    boxes, stride = L.as_f32_rows(boxes)
    scores = scores.float().contiguous()
    K = boxes.size(0)
    dev = boxes.device
    keep = torch.empty((K,), dtype=torch.int64, device=dev)
    if K == 0:
        return keep, torch.zeros((n_batches,) if batch_ids is not None else (), dtype=torch.int64, device=dev)
    num = torch.empty((n_batches,) if batch_ids is not None else (), dtype=torch.int64, device=dev)   # zeroed by the library
    if batch_ids is not None:
        L.require_cuda(batch_ids)
        batch_ids = batch_ids.to(torch.int64).contiguous()
    if labels is not None:
        L.require_cuda(labels)
        labels = labels.to(torch.int64).contiguous()
    if class_offset is not None:
        class_offset = class_offset.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if class_offset.numel() != (n_batches if batch_ids is not None else 1):
            raise ValueError('class_offset must hold one scale per image')
    flags = (L.NMS_INCLUSIVE if inclusive else 0) | (L.NMS_ORDER_INDEX if order_index else 0) | \
            (L.NMS_DROP_SMALL if drop_small else 0) | (L.NMS_STRICT if strict else 0) | (L.NMS_SORT_PATH if sort_path else 0) | \
            ((int(label_bits) & 63) << 8)
    lib = L.lib()
    nbytes = C.c_size_t(0)
    L.check(lib.r3g_nms_workspace_bytes(K, C.byref(nbytes)))
    ws = L.workspace(nbytes.value, dev)
    with L.device_guard(dev):
        L.check(lib.r3g_nms_batched_counted_f32(L.ptr(boxes), stride, L.ptr(scores), L.ptr(labels), L.ptr(batch_ids),
                                                int(n_batches), K, None if count is None else C.c_void_p(count.data_ptr()), float(thr),
                                                L.V[variant], flags, L.ptr(class_offset), L.ptr(keep), C.c_void_p(num.data_ptr()),
                                                L.ptr(ws), ws.numel(), L.stream_ptr(dev)))
    if stats is not None:
        stats['counters'] = ws[512:1024].view(torch.int64)
    return keep, num


def pack_keep_records(boxes, scores, labels, keep, num, batch_ids=None, n_batches=1, max_per_img=2000, drop_last=False):
    """Fixed-size detections from the outputs of `nms_device`, with no host synchronisation: returns
    (dets (n_batches, max_per_img, 6), labels (n_batches, max_per_img) int64, counts (n_batches,) int64); image b's rows
    [0, counts[b]) are its kept <x, y, w, h, a, score> in keep order (what `dets[keep][:max_per_img]` gives), the rest zero."""
    L.require_cuda(boxes, scores, keep, num)
    boxes, stride = L.as_f32_rows(boxes)
    scores = scores.float().contiguous()
    dev = boxes.device
    K = boxes.size(0)
    if labels is not None:
        labels = labels.to(torch.int64).contiguous()
    if batch_ids is not None:
        batch_ids = batch_ids.to(torch.int64).contiguous()
    num = num.reshape(-1).to(torch.int64).contiguous()
    dets = torch.empty((n_batches, max_per_img, 6), dtype=torch.float32, device=dev)
    labs = torch.empty((n_batches, max_per_img), dtype=torch.int64, device=dev)
    counts = torch.empty((n_batches,), dtype=torch.int64, device=dev)
    with L.device_guard(dev):
        L.check(L.lib().r3g_nms_pack_f32(L.ptr(boxes), stride, L.ptr(scores), L.ptr(labels), L.ptr(keep), C.c_void_p(num.data_ptr()),
                                         L.ptr(batch_ids), int(n_batches), K, int(max_per_img), int(bool(drop_last)),
                                         C.c_void_p(dets.data_ptr()), C.c_void_p(labs.data_ptr()), C.c_void_p(counts.data_ptr()),
                                         L.stream_ptr(dev)))
    return dets, labs, counts


def to_cuda_input(x, device_id, what):
    """numpy / CPU tensors are uploaded (the reference's convenience path, e.g. rnms_wrapper.py:11-15);
    returns (cuda tensor, is_numpy, was_host).  Host inputs get the reference's CPU rule (IoU >= thr)."""
    import numpy as np
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x, False, False
        dev = torch.device("cuda", torch.cuda.current_device() if device_id is None else device_id)
        return x.to(dev), False, True
    if isinstance(x, np.ndarray):
        dev = torch.device("cuda", torch.cuda.current_device() if device_id is None else device_id)
        return torch.from_numpy(x).to(dev), True, device_id is None
    raise TypeError(f"{what} must be either a Tensor or numpy array, but got {type(x)}")
