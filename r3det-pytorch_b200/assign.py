"""Fused max-IoU assignment — the consumer side of the IoU calculators.

In the reference, mmdet-2.19's ``MaxIoUAssigner`` calls ``RBboxOverlaps2D_v*`` for the full (G, A) matrix
(800 MB for 1000 x 200k) and then reduces it with ``max`` over both axes
(caller: r3det/models/dense_heads/rotate_anchor_head.py:220-228; assign_wrt_overlaps semantics recalled in SURVEY.md A6).
``max_iou_assign`` does the same assignment in one C call without materialising the matrix; the overlaps it reduces
are the matrix kernel's values bit for bit.  ``FusedMaxIoUAssigner`` wraps it behind MaxIoUAssigner's constructor and
``assign`` signature and registers itself in mmdet's BBOX_ASSIGNERS when mmdet is importable: tuple ``neg_iou_thr``, ignore
regions (``ignore_iof_thr`` > 0 with ``gt_bboxes_ignore``, both ``ignore_wrt_candidates`` settings) and
``gt_max_assign_all`` / ``match_low_quality`` behave as in mmdet; ``gpu_assign_thr`` (mmdet's CPU off-load for many GTs) is
accepted and ignored — there is no CPU path here, and the fused kernel never holds the (G, A) matrix it exists to avoid."""
import ctypes as C
from collections import namedtuple

import torch

from . import _lib as L

AssignOutput = namedtuple('AssignOutput', 'num_gts gt_inds max_overlaps labels argmax_overlaps gt_max_overlaps gt_argmax_overlaps')


def max_iou_assign(gt_bboxes, bboxes, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, match_low_quality=True,
                   gt_max_assign_all=True, variant='v1', gt_labels=None, flags=L.FLAG_STRICT):
    """Assign each of the A `bboxes` (anchors) to a GT index + 1, 0 (background) or -1 (ignore).

    Returns AssignOutput(num_gts, gt_inds (A,) int64, max_overlaps (A,), labels (A,) or None,
    argmax_overlaps (A,), gt_max_overlaps (G,), gt_argmax_overlaps (G,))."""
    L.require_cuda(bboxes)
    anchors, sa = L.as_f32_rows(bboxes)
    A = anchors.size(0)
    dev = anchors.device
    if gt_bboxes is None or gt_bboxes.numel() == 0:
        gt, sg, G = None, 5, 0
    else:
        L.require_cuda(gt_bboxes)
        gt, sg = L.as_f32_rows(gt_bboxes)
        G = gt.size(0)
    assigned = torch.empty((A,), dtype=torch.int64, device=dev)
    max_ov = torch.empty((A,), dtype=torch.float32, device=dev)
    argmax = torch.empty((A,), dtype=torch.int64, device=dev)
    gt_max = torch.empty((G,), dtype=torch.float32, device=dev)
    gt_arg = torch.empty((G,), dtype=torch.int64, device=dev)
    if A:
        lib = L.lib()
        need = C.c_size_t(0)
        L.check(lib.r3g_assign_workspace_bytes(G, A, C.byref(need)))
        ws = L.workspace(need.value, dev)
        with L.device_guard(dev):
            L.check(lib.r3g_max_iou_assign_f32(L.ptr(gt), G, sg, L.ptr(anchors), A, sa, L.V[variant], flags,
                                               float(pos_iou_thr), float(neg_iou_thr), float(min_pos_iou),
                                               int(bool(match_low_quality)), int(bool(gt_max_assign_all)),
                                               L.ptr(assigned), L.ptr(max_ov), L.ptr(argmax), L.ptr(gt_max), L.ptr(gt_arg),
                                               L.ptr(ws), ws.numel(), L.stream_ptr(dev)))
    labels = None
    if gt_labels is not None:
        labels = assigned.new_full((A,), -1)
        pos = assigned > 0
        if G:
            labels[pos] = gt_labels.to(dev)[assigned[pos] - 1]
    return AssignOutput(G, assigned, max_ov, labels, argmax, gt_max, gt_arg)


def max_iou_assign_batched(gt_bboxes_list, bboxes, pos_iou_thr, neg_iou_thr, min_pos_iou=0.0, match_low_quality=True,
                           gt_max_assign_all=True, variant='v1', flags=L.FLAG_STRICT):
    """`max_iou_assign` for a batch of images in one launch sequence (what the anchor heads do per image in a Python loop,
    rotate_anchor_head.py:316-333 via multi_apply).  gt_bboxes_list: B tensors (G_b, >=5), empty ones allowed;
    bboxes: (A, >=5) anchors shared by the batch, or (B, A, >=5) per-image boxes (refine stage).
    Returns a list of B AssignOutput (labels None); the per-anchor tensors are rows of one (B, A) allocation."""
    B = len(gt_bboxes_list)
    L.require_cuda(bboxes)
    shared = bboxes.dim() == 2
    if not shared:
        assert bboxes.dim() == 3 and bboxes.size(0) == B
    if B > 64:
        out = []
        for s in range(0, B, 64):
            out += max_iou_assign_batched(gt_bboxes_list[s:s + 64], bboxes if shared else bboxes[s:s + 64], pos_iou_thr, neg_iou_thr,
                                          min_pos_iou, match_low_quality, gt_max_assign_all, variant, flags)
        return out
    anchors = bboxes.float().contiguous()
    A, sa = anchors.size(-2), anchors.size(-1)
    assert sa >= 5
    dev = anchors.device
    counts = [0 if g is None else int(g.size(0)) for g in gt_bboxes_list]
    NR = sum(counts)
    if NR:
        parts = [g.float()[:, :5] for g, c in zip(gt_bboxes_list, counts) if c]
        for g in parts:
            L.require_cuda(g)
        gt = torch.cat(parts).contiguous()
    else:
        gt = None
    assigned = torch.empty((B, A), dtype=torch.int64, device=dev)
    max_ov = torch.empty((B, A), dtype=torch.float32, device=dev)
    argmax = torch.empty((B, A), dtype=torch.int64, device=dev)
    gt_max = torch.empty((NR,), dtype=torch.float32, device=dev)
    gt_arg = torch.empty((NR,), dtype=torch.int64, device=dev)
    if A and B:
        lib = L.lib()
        cnt = (C.c_int64 * B)(*counts)
        need = C.c_size_t(0)
        L.check(lib.r3g_assign_batched_workspace_bytes(B, cnt, A, int(shared), C.byref(need)))
        ws = L.workspace(need.value, dev)
        with L.device_guard(dev):
            L.check(lib.r3g_max_iou_assign_batched_f32(B, L.ptr(gt), cnt, 5, L.ptr(anchors), A, sa, int(shared), L.V[variant], flags,
                                                       float(pos_iou_thr), float(neg_iou_thr), float(min_pos_iou),
                                                       int(bool(match_low_quality)), int(bool(gt_max_assign_all)),
                                                       L.ptr(assigned), L.ptr(max_ov), L.ptr(argmax), L.ptr(gt_max), L.ptr(gt_arg),
                                                       L.ptr(ws), ws.numel(), L.stream_ptr(dev)))
    out, r0 = [], 0
    for b in range(B):
        out.append(AssignOutput(counts[b], assigned[b], max_ov[b], None, argmax[b], gt_max[r0:r0 + counts[b]], gt_arg[r0:r0 + counts[b]]))
        r0 += counts[b]
    return out


class FusedMaxIoUAssigner(object):
    """MaxIoUAssigner with the IoU calculator fused in (same constructor arguments; `iou_calculator` selects the variant)."""

    def __init__(self, pos_iou_thr, neg_iou_thr, min_pos_iou=.0, gt_max_assign_all=True, ignore_iof_thr=-1,
                 ignore_wrt_candidates=True, match_low_quality=True, gpu_assign_thr=-1,
                 iou_calculator=dict(type='RBboxOverlaps2D_v1')):
        if isinstance(neg_iou_thr, (tuple, list)):
            assert len(neg_iou_thr) == 2
            neg_iou_thr = (float(neg_iou_thr[0]), float(neg_iou_thr[1]))
        kind = iou_calculator['type'] if isinstance(iou_calculator, dict) else type(iou_calculator).__name__
        self.variant = {'RBboxOverlaps2D_v1': 'v1', 'RBboxOverlaps2D_v2': 'v2', 'RBboxOverlaps2D_v3': 'v3'}[kind]
        self.pos_iou_thr, self.neg_iou_thr, self.min_pos_iou = pos_iou_thr, neg_iou_thr, min_pos_iou
        self.gt_max_assign_all, self.match_low_quality = gt_max_assign_all, match_low_quality
        self.ignore_iof_thr, self.ignore_wrt_candidates = ignore_iof_thr, ignore_wrt_candidates
        self.gpu_assign_thr = gpu_assign_thr          # accepted for config compatibility; no CPU off-load exists here

    def _ignored(self, bboxes, gt_bboxes_ignore, flags):
        """Anchors whose best IoF with an ignore region exceeds ignore_iof_thr (mmdet sets their overlap column to -1)."""
        from .rbbox_geo import pairwise_iou
        if self.ignore_wrt_candidates:
            iof = pairwise_iou(bboxes, gt_bboxes_ignore, self.variant, 'iof', flags).max(dim=1).values
        else:
            iof = pairwise_iou(gt_bboxes_ignore, bboxes, self.variant, 'iof', flags).max(dim=0).values
        return iof > self.ignore_iof_thr

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None):
        strip = lambda b: b[..., :5] if b is not None and b.size(-1) == 6 else b
        bboxes, gt_bboxes, gt_bboxes_ignore = strip(bboxes), strip(gt_bboxes), strip(gt_bboxes_ignore)
        flags = L.FLAG_STRICT | (L.FLAG_SMALL_MASK if self.variant == 'v3' else 0)
        tuple_neg = isinstance(self.neg_iou_thr, tuple)
        neg = self.neg_iou_thr[1] if tuple_neg else self.neg_iou_thr
        args = (self.pos_iou_thr, neg, self.min_pos_iou, self.match_low_quality, self.gt_max_assign_all, self.variant)
        keep = None
        if (self.ignore_iof_thr > 0 and gt_bboxes_ignore is not None and gt_bboxes_ignore.numel() > 0 and bboxes.numel() > 0
                and gt_bboxes is not None and gt_bboxes.numel() > 0):
            ign = self._ignored(bboxes, gt_bboxes_ignore, flags)
            if bool(ign.any()):
                keep = (~ign).nonzero(as_tuple=False).squeeze(1)
        if keep is None:
            out = max_iou_assign(gt_bboxes, bboxes, *args, gt_labels=None, flags=flags)
        else:
            # ignored anchors never enter the sweep: their column is -1 for every GT, so they cannot be a maximum of anything
            sub = max_iou_assign(gt_bboxes, bboxes[keep], *args, gt_labels=None, flags=flags)
            A = bboxes.size(0)
            gt_inds = sub.gt_inds.new_full((A,), -1); gt_inds[keep] = sub.gt_inds
            max_ov = sub.max_overlaps.new_full((A,), -1.0); max_ov[keep] = sub.max_overlaps
            argmax = sub.argmax_overlaps.new_zeros((A,)); argmax[keep] = sub.argmax_overlaps
            gt_arg = keep[sub.gt_argmax_overlaps] if keep.numel() else sub.gt_argmax_overlaps
            out = AssignOutput(sub.num_gts, gt_inds, max_ov, None, argmax, sub.gt_max_overlaps, gt_arg)
        gt_inds = out.gt_inds
        if tuple_neg and out.num_gts > 0:
            # background only inside [neg_lo, neg_hi): what the kernel marked 0 below neg_lo goes back to -1
            gt_inds = torch.where((gt_inds == 0) & (out.max_overlaps < self.neg_iou_thr[0]), gt_inds.new_full((), -1), gt_inds)
        labels = None
        if gt_labels is not None:
            labels = gt_inds.new_full((gt_inds.numel(),), -1)
            pos = gt_inds > 0
            if out.num_gts:
                labels[pos] = gt_labels.to(gt_inds.device)[gt_inds[pos] - 1]
        out = AssignOutput(out.num_gts, gt_inds, out.max_overlaps, labels, out.argmax_overlaps, out.gt_max_overlaps,
                           out.gt_argmax_overlaps)
        try:  # pragma: no cover - mmdet is not installed in the build image
            from mmdet.core.bbox.assigners import AssignResult
            return AssignResult(out.num_gts, out.gt_inds, out.max_overlaps, labels=out.labels)
        except Exception:  # noqa: BLE001
            return out


try:  # pragma: no cover
    from mmdet.core.bbox.builder import BBOX_ASSIGNERS
    BBOX_ASSIGNERS.register_module(name='FusedMaxIoUAssigner', force=True)(FusedMaxIoUAssigner)
except Exception:  # noqa: BLE001
    pass
