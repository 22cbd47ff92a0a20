"""v3 rotated IoU — host-side mirror of r3det/ops/box_iou_rotated/box_iou_rotated_wrapper.py:8-64.

`obb_overlaps(bboxes1, bboxes2, mode, is_aligned, device_id)`: matrix mode goes through the IoU kernel with
the wrapper's too-small mask (rows/cols with min(w,h) < 1e-3 -> 0, :54-60); empty inputs give zeros (:43-46).
Aligned mode returns (m, 1); it reuses the matrix geometry core — the reference's pure-torch aligned path
disagrees with its own matrix op by up to 2.7e-4, so parity is against the matrix op.  Like that torch
path (aligned_obb_overlaps, :67-92) it applies NO too-small mask; unlike it, it is not differentiable: inputs that require
grad raise (no shipped config differentiates through it).
numpy inputs are uploaded to cuda:device_id (current device if None) and returned as numpy."""
import numpy as np
import torch

from . import _lib as L
from .rbbox_geo import aligned_iou, pairwise_iou


def obb_overlaps(bboxes1, bboxes2, mode='iou', is_aligned=False, device_id=None):
    assert mode in ['iou', 'iof']
    assert type(bboxes1) is type(bboxes2)
    if is_aligned:
        assert bboxes1.shape[0] == bboxes2.shape[0]
    if isinstance(bboxes1, torch.Tensor):
        is_numpy = False
        b1, b2 = bboxes1, bboxes2
        if not b1.is_cuda:
            dev = torch.device('cuda', torch.cuda.current_device())
            b1, b2 = b1.to(dev), b2.to(dev)
    elif isinstance(bboxes1, np.ndarray):
        is_numpy = True
        dev = torch.device('cuda', torch.cuda.current_device() if device_id is None else device_id)
        b1 = torch.from_numpy(bboxes1).float().to(dev)
        b2 = torch.from_numpy(bboxes2).float().to(dev)
    else:
        raise TypeError('bboxes must be either a Tensor or numpy array, '
                        f'but got {type(bboxes1)}')
    if b1.numel() == 0 or b2.numel() == 0:
        rows, cols = b1.size(0), b2.size(0)
        outputs = b1.new_zeros(rows, 1) if is_aligned else b1.new_zeros(rows, cols)
    elif is_aligned:
        L.require_no_grad('obb_overlaps(is_aligned=True)', b1, b2)
        outputs = aligned_iou(b1, b2, 'v3', mode, L.FLAG_STRICT)[:, None]
    else:
        outputs = pairwise_iou(b1, b2, 'v3', mode, L.FLAG_STRICT | L.FLAG_SMALL_MASK)
    if isinstance(bboxes1, torch.Tensor) and bboxes1.is_floating_point() and outputs.dtype != bboxes1.dtype:
        outputs = outputs.to(bboxes1.dtype)                    # same dtype out as in, like the reference op; FP32 arithmetic
    if is_numpy:
        outputs = outputs.cpu().numpy()
    elif isinstance(bboxes1, torch.Tensor) and not bboxes1.is_cuda:
        outputs = outputs.cpu()
    return outputs
