"""multiclass_nms_rotated — same call surface and results as
r3det/core/post_processing/bbox_nms_rotated.py:7-131, different machinery.

The reference expands the boxes to (n, C, 5), boolean-masks them, calls `nonzero` (host sync) and then one of four
NMS back ends.  Here one CUDA pass (`r3g_mc_candidates_f32`) emits the candidate list in the same row-major
(box, class) order, the per-class segmented NMS kernel does the rest, and a small table carries what differs between
the four `nms.type` values:

  type    geometry  class offset scale (reference wrapper)           keep order        after NMS
  'v1'    v1        bboxes.max() + 1                (rnms_wrapper.py:61-64)   candidate index   dets[:max_num]
  'v3'    v3        hbb.max() - hbb.min() + 1       (nms_rotated_wrapper.py:84-90)  score      dets[:max_num]
  'v2'    v2        none (labels compared)          (ml_nms_rotated)          score             re-sort + slice if k > max_num
  'mmcv'  v2        none (mmcv.ops.nms_rotated with labels; third-party)      score             dets[:max_num], return_inds

Quirks kept on purpose: 'v1' truncates in candidate-index order; 'v2' with the default max_num=-1 drops the
lowest-scoring detection (`keep.size(0) > -1` → `inds[:-1]`, :119-121); `score_factors` are applied after the
threshold test; empty input returns ((0, 6), (0,)) int64 labels."""
import ctypes as C

import torch

from . import _lib as L
from ._nms_core import nms_device, pack_keep_records
from .nms_rotated import obb2hbb as _obb2xyxy_v3

#            geometry, ordered by index, drop tiny boxes, offset rule
_SPEC = {
    'v1':   ('v1', True,  False, 'max'),
    'v3':   ('v3', False, True,  'hbb_span'),
    'v2':   ('v2', False, False, None),
    'mmcv': ('v2', False, False, None),
}


def _cfg(nms, key, default=None):
    return nms.get(key, default) if isinstance(nms, dict) else getattr(nms, key, default)


def _candidates(multi_bboxes, multi_scores, score_thr, score_factors):
    """Device-side candidate list: boxes (K,5), scores (K,), labels (K,), flat (box*C + class) index (K,)."""
    n, C1 = multi_scores.shape
    nc = C1 - 1
    dev = multi_scores.device
    mb = multi_bboxes.float().contiguous()
    ms = multi_scores.float().contiguous()
    sf = None if score_factors is None else score_factors.float().contiguous()
    T = n * nc
    boxes = torch.empty((T, 5), dtype=torch.float32, device=dev)
    scores = torch.empty((T,), dtype=torch.float32, device=dev)
    labels = torch.empty((T,), dtype=torch.int64, device=dev)
    src = torch.empty((T,), dtype=torch.int64, device=dev)
    count = torch.zeros((), dtype=torch.int64, device=dev)
    if T:
        lib = L.lib()
        nbytes = C.c_size_t(0)
        L.check(lib.r3g_mc_candidates_workspace_bytes(n, nc, C.byref(nbytes)))
        ws = L.workspace(nbytes.value, dev)
        with L.device_guard(dev):
            L.check(lib.r3g_mc_candidates_f32(L.ptr(mb), mb.size(1), L.ptr(ms), C1, L.ptr(sf), n, nc, float(score_thr),
                                              L.ptr(boxes), L.ptr(scores), L.ptr(labels), L.ptr(src),
                                              C.c_void_p(count.data_ptr()), L.ptr(ws), ws.numel(), L.stream_ptr(dev)))
    K = int(count.item())
    return boxes[:K], scores[:K], labels[:K], src[:K]


def multiclass_nms_rotated(multi_bboxes, multi_scores, score_thr, nms, max_num=-1, score_factors=None,
                           return_inds=False):
    """NMS for multi-class rotated bboxes: (dets (k, 6), labels (k,)[, flat indices (k,)])."""
    kind = _cfg(nms, 'type', 'v1')
    if kind not in _SPEC:
        raise KeyError(f'unknown rotated nms type {kind!r}')
    geometry, by_index, drop_small, offset_rule = _SPEC[kind]
    iou_thr = _cfg(nms, 'iou_thr')
    host_in = not multi_scores.is_cuda                  # CPU tensors: upload, reference CPU rule (>=), results back on CPU
    dev = multi_scores.device if not host_in else torch.device('cuda', torch.cuda.current_device())
    back = (lambda t: t.cpu()) if host_in else (lambda t: t)

    boxes, scores, labels, src = _candidates(multi_bboxes.to(dev), multi_scores.to(dev), score_thr,
                                             None if score_factors is None else score_factors.to(dev))
    if boxes.size(0) == 0:
        if kind == 'mmcv':
            dets = torch.cat([boxes, scores[:, None]], -1)
            return (back(dets), back(labels), back(src)) if return_inds else (back(dets), back(labels))
        return multi_bboxes.new_zeros((0, 6)), multi_bboxes.new_zeros((0, ), dtype=torch.long)

    scale = None
    if offset_rule == 'max':
        scale = boxes.max() + 1
    elif offset_rule == 'hbb_span':
        hb = _obb2xyxy_v3(boxes)
        scale = (hb.max() - hb.min()) + 1
    keep, num = nms_device(boxes, scores, iou_thr, geometry, labels=labels, class_offset=scale, inclusive=host_in,
                           order_index=by_index, drop_small=drop_small, label_bits=max(1, (multi_scores.size(1) - 2).bit_length()))
    keep = keep[:int(num.item())]

    if kind == 'v2' and keep.size(0) > max_num:          # reference :119-124 (also fires for max_num = -1)
        keep = keep[:max_num]                            # keep is already in descending-score order
    elif kind != 'v2' and max_num > 0:
        keep = keep[:max_num]
    dets = torch.cat([boxes[keep], scores[keep][:, None]], 1)
    if kind == 'mmcv' and return_inds:
        return back(dets), back(labels[keep]), back(keep)
    return back(dets), back(labels[keep])


def multiclass_nms_rotated_batch(multi_bboxes, multi_scores, score_thr, nms, max_num=-1):
    """`multiclass_nms_rotated` for a whole batch in one launch sequence: multi_bboxes (B, n, 5), multi_scores
    (B, n, C + 1) CUDA tensors -> list of B (dets (k, 6), labels (k,)) tuples, each identical to what the per-image
    call returns.  With max_num > 0 (every reference config) the batch runs on the fixed-size device path of
    `multiclass_nms_rotated_padded` and the host reads the B keep counts ONCE, after the detections exist; with
    max_num <= 0 there are two host synchronisations per BATCH (candidate count, keep counts).  The reference loops
    over images (rotate_anchor_head.py:565-588) with ~40 launches and a `nonzero` sync each."""
    kind = _cfg(nms, 'type', 'v1')
    if kind not in _SPEC:
        raise KeyError(f'unknown rotated nms type {kind!r}')
    geometry, by_index, drop_small, offset_rule = _SPEC[kind]
    L.require_cuda(multi_bboxes, multi_scores)
    assert multi_bboxes.dim() == 3 and multi_scores.dim() == 3 and multi_bboxes.size(-1) == 5
    B, n, C1 = multi_scores.shape
    nc = C1 - 1
    empty = (multi_bboxes.new_zeros((0, 6)), multi_bboxes.new_zeros((0,), dtype=torch.long))
    if B == 0:
        return []
    if max_num > 0 and n * nc > 0:
        # fixed-size device path (no host read before the detections exist), then ONE read of the B counts to cut the views
        dets, labels, counts = multiclass_nms_rotated_padded(multi_bboxes, multi_scores, score_thr, nms, max_num)
        return [(dets[b, :c], labels[b, :c]) for b, c in enumerate(counts.tolist())]
    boxes, scores, labels, src = _candidates(multi_bboxes.reshape(B * n, 5), multi_scores.reshape(B * n, C1), score_thr, None)
    if boxes.size(0) == 0:
        return [empty for _ in range(B)]
    bid = torch.div(src, n * nc, rounding_mode='floor')
    scale = None
    if offset_rule is not None:
        # per-image offset scale over the image's CANDIDATE boxes = rows with at least one class above the threshold;
        # dense masked reductions over (B, n) instead of a scatter with B hot addresses
        rowmask = (multi_scores[..., :nc] > score_thr).any(-1)
        mb = multi_bboxes.float()
        if offset_rule == 'max':
            top = mb.max(-1).values.masked_fill(~rowmask, float('-inf')).max(-1).values
            scale = top + 1
        else:
            hb = _obb2xyxy_v3(mb.reshape(B * n, 5)).reshape(B, n, 4)
            top = hb.max(-1).values.masked_fill(~rowmask, float('-inf')).max(-1).values
            bot = hb.min(-1).values.masked_fill(~rowmask, float('inf')).min(-1).values
            scale = (top - bot) + 1
        scale = torch.where(torch.isfinite(scale), scale, torch.ones_like(scale))      # images without candidates
    keep, num = nms_device(boxes, scores, _cfg(nms, 'iou_thr'), geometry, labels=labels, class_offset=scale,
                           order_index=by_index, drop_small=drop_small, batch_ids=bid, n_batches=B,
                           label_bits=max(1, (nc - 1).bit_length()))
    counts = num.tolist()
    total = sum(counts)
    keep = keep[:total]
    dets_kept = torch.cat([boxes[keep], scores[keep][:, None]], 1)          # one gather for the whole batch, views per image
    labels_kept = labels[keep]
    out, start = [], 0
    for b in range(B):
        c = counts[b]
        if c == 0:
            out.append(empty)
            continue
        m = c
        if kind == 'v2' and c > max_num:
            m = len(range(c)[:max_num])                                     # the reference's keep[:max_num], also for max_num = -1
        elif kind != 'v2' and max_num > 0:
            m = min(c, max_num)
        out.append((dets_kept[start:start + m], labels_kept[start:start + m]))
        start += c
    return out


_RULE = {None: 0, 'max': 1, 'hbb_span': 2}


def multiclass_nms_rotated_padded(multi_bboxes, multi_scores, score_thr, nms, max_num):
    """`multiclass_nms_rotated` for a batch with FIXED-SIZE outputs and no host synchronisation anywhere (CUDA-graph
    capturable): multi_bboxes (B, n, 5), multi_scores (B, n, C + 1) -> dets (B, max_num, 6), labels (B, max_num) int64,
    counts (B,) int64.  Rows [0, counts[b]) of image b are exactly what the per-image call returns (same order and
    truncation, reference bbox_nms_rotated.py:98-131); the remaining rows are zero.  Candidate extraction, the per-image
    class-offset scale, the segmented NMS and the truncation all run on the device: r3g_mc_candidates_batched_f32 ->
    r3g_nms_batched_counted_f32 -> r3g_nms_pack_f32.  max_num > 0; large batches are processed in slices (see below)."""
    kind = _cfg(nms, 'type', 'v1')
    if kind not in _SPEC:
        raise KeyError(f'unknown rotated nms type {kind!r}')
    geometry, by_index, drop_small, offset_rule = _SPEC[kind]
    L.require_cuda(multi_bboxes, multi_scores)
    assert multi_bboxes.dim() == 3 and multi_scores.dim() == 3 and multi_bboxes.size(-1) == 5 and max_num > 0
    B, n, C1 = multi_scores.shape
    nc = C1 - 1
    dev = multi_scores.device
    # the candidate buffers and the NMS workspace (~0.5 KB per slot) are sized for the CAPACITY B * n * C of the call, not for
    # the candidates that exist: batches are cut so that one call stays at <= 2^21 slots (~1 GB of workspace) and <= 64 images
    step = min(64, max(1, (1 << 21) // max(n * nc, 1)))
    if B > step:
        parts = [multiclass_nms_rotated_padded(multi_bboxes[s:s + step], multi_scores[s:s + step], score_thr, nms, max_num)
                 for s in range(0, B, step)]
        return tuple(torch.cat([p[i] for p in parts]) for i in range(3))
    T = B * n * nc
    if T == 0:
        return (multi_bboxes.new_zeros((B, max_num, 6)), torch.zeros((B, max_num), dtype=torch.int64, device=dev),
                torch.zeros((B,), dtype=torch.int64, device=dev))
    mb = multi_bboxes.float().contiguous()
    ms = multi_scores.float().contiguous()
    boxes = torch.empty((T, 5), dtype=torch.float32, device=dev)
    scores = torch.empty((T,), dtype=torch.float32, device=dev)
    labels = torch.empty((T,), dtype=torch.int64, device=dev)
    bid = torch.empty((T,), dtype=torch.int64, device=dev)
    src = torch.empty((T,), dtype=torch.int64, device=dev)
    count = torch.empty((), dtype=torch.int64, device=dev)
    scale = torch.empty((B,), dtype=torch.float32, device=dev)
    lib = L.lib()
    nbytes = C.c_size_t(0)
    L.check(lib.r3g_mc_candidates_workspace_bytes(B * n, nc, C.byref(nbytes)))
    ws = L.workspace(nbytes.value + 512, dev)
    with L.device_guard(dev):
        L.check(lib.r3g_mc_candidates_batched_f32(L.ptr(mb), 5, L.ptr(ms), C1, n, B, nc, float(score_thr), _RULE[offset_rule],
                                                  L.ptr(boxes), L.ptr(scores), L.ptr(labels), L.ptr(bid), L.ptr(src),
                                                  C.c_void_p(count.data_ptr()), L.ptr(scale), L.ptr(ws), ws.numel(),
                                                  L.stream_ptr(dev)))
    keep, num = nms_device(boxes, scores, _cfg(nms, 'iou_thr'), geometry, labels=labels,
                           class_offset=scale if offset_rule is not None else None, order_index=by_index, drop_small=drop_small,
                           batch_ids=bid, n_batches=B, label_bits=max(1, (nc - 1).bit_length()), count=count)
    return pack_keep_records(boxes, scores, labels, keep, num, bid, B, int(max_num))
