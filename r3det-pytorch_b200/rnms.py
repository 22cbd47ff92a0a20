"""v1 rotated NMS — host-side mirror of r3det/ops/rnms/rnms_wrapper.py (rnms :7-31, batched_rnms :34-69).

Same signatures and results as the reference: `rnms` returns kept detections and their ORIGINAL indices
sorted ascending (rnms_kernel.cu:331-334 / rnms_cpu.cpp:281).  CUDA inputs follow the reference GPU rule
(IoU > thr, rnms_kernel.cu:260); numpy / CPU inputs are uploaded and follow the reference CPU rule
(IoU >= thr, rnms_cpu.cpp:277).  All compute is on the GPU.
"""
import torch

from ._nms_core import nms_device, to_cuda_input


def rnms(dets, nms_thr, device_id=None):
    """Compute NMS of oriented bboxes.  dets: (K, 6) <x, y, w, h, a, score>."""
    dets_th, is_numpy, was_host = to_cuda_input(dets, device_id, "dets")
    if dets_th.shape[0] == 0:
        inds = dets_th.new_zeros(0, dtype=torch.long)
    else:
        d = dets_th.float()
        keep, num = nms_device(d[:, :5], d[:, 5], nms_thr, "v1", inclusive=was_host, order_index=True)
        inds = keep[:int(num.item())]
    if is_numpy:
        inds = inds.cpu().numpy()
    elif was_host:
        inds = inds.cpu()
    return dets[inds, :], inds


def batched_rnms(bboxes, scores, inds, nms_thr, class_agnostic=False):
    """NMS per cluster id `inds` (class).  The reference offsets x,y by inds*(bboxes.max()+1) and runs one
    K x K NMS (rnms_wrapper.py:58-66); here classes are segmented on the device and the same FP32 offset is
    applied inside the kernel so the geometry is evaluated on identical coordinates."""
    if class_agnostic or bboxes.shape[0] == 0:
        dets, keep = rnms(torch.cat([bboxes, scores[:, None]], -1), nms_thr)
        return torch.cat([bboxes[keep], dets[:, -1:]], -1), keep
    b, _, was_host = to_cuda_input(bboxes, None, "bboxes")
    s = scores.to(b.device)
    lab = inds.to(b.device)
    scale = b.max() + 1
    keep, num = nms_device(b, s, nms_thr, "v1", labels=lab, class_offset=scale, inclusive=was_host, order_index=True)
    keep = keep[:int(num.item())]
    if was_host:
        keep = keep.cpu()
    bboxes = bboxes[keep]
    scores = scores[keep]
    return torch.cat([bboxes, scores[:, None]], -1), keep
