"""Box transforms — mirror of the torch half of r3det/core/bbox/rtransforms.py (dispatchers :49-187).

poly2obb / obb2poly / obb2hbb / hbb2obb / obb2xyxy take the reference's (tensor, version) arguments and run
one CUDA kernel each (csrc/transforms.cu) instead of 20-40 small torch kernels.  CUDA tensors only.
rbbox2result / rbbox2roi / norm_angle are small host/torch helpers kept for API completeness."""
import numpy as np
import torch

from . import _lib as L

_VER = {'v1': 1, 'v2': 2, 'v3': 3}


def _run(fn_name, x, version, in_cols, out_cols):
    if version not in _VER:
        raise NotImplementedError
    L.require_cuda(x)
    lead = x.shape[:-1]
    xin = x.reshape(-1, in_cols)
    if xin.dtype != torch.float32:
        xin = xin.float()
    xin = xin.contiguous()
    n = xin.size(0)
    out = torch.empty((n, out_cols), dtype=torch.float32, device=x.device)
    if n:
        with L.device_guard(x.device):
            L.check(getattr(L.lib(), fn_name)(L.ptr(xin), n, _VER[version], L.ptr(out), L.stream_ptr(x.device)))
    return out.reshape(*lead, out_cols).to(x.dtype) if x.dtype != torch.float32 else out.reshape(*lead, out_cols)


def poly2obb(polys, version='v1'):
    """[x0,y0,...,x3,y3] -> [x_ctr,y_ctr,w,h,angle] (rtransforms.py:49-67); input is reshaped to (-1, 8)."""
    return _run('r3g_poly2obb_f32', polys.reshape(-1, 8), version, 8, 5)


def obb2poly(rbboxes, version='v1'):
    """[x_ctr,y_ctr,w,h,angle] -> [x0,y0,...,x3,y3] (rtransforms.py:110-127)."""
    return _run('r3g_obb2poly_f32', rbboxes[..., :5], version, 5, 8)


def obb2hbb(rbboxes, version='v1'):
    """oriented -> horizontal box in obb format (rtransforms.py:91-107)."""
    return _run('r3g_obb2hbb_f32', rbboxes, version, 5, 5)


def hbb2obb(hbboxes, version='v1'):
    """[x_lt,y_lt,x_rb,y_rb] -> [x_ctr,y_ctr,w,h,angle] (rtransforms.py:168-187).
    As in the reference, v1 slices the last dimension with 0::4, so (N, 4k) input gives (N, k, 5) — (N, 4)
    gives (N, 1, 5) (rtransforms.py:548-554); v2/v3 map (..., 4) to (..., 5)."""
    if version == 'v1':
        n = hbboxes.size(0)
        return _run('r3g_hbb2obb_f32', hbboxes.reshape(n, -1, 4), version, 4, 5)
    return _run('r3g_hbb2obb_f32', hbboxes, version, 4, 5)


def obb2xyxy(rbboxes, version='v1'):
    """oriented -> [x_lt,y_lt,x_rb,y_rb] (rtransforms.py:149-165)."""
    return _run('r3g_obb2xyxy_f32', rbboxes, version, 5, 4)


def norm_angle(angle, angle_range):
    """Limit the range of angles (rtransforms.py:789-805)."""
    if angle_range == 'v1':
        return angle
    elif angle_range == 'v2':
        return (angle + np.pi / 4) % np.pi - np.pi / 4
    elif angle_range == 'v3':
        return (angle + np.pi / 2) % np.pi - np.pi / 2
    else:
        print('Not yet implemented.')


def rbbox2result(bboxes, labels, num_classes):
    """Detections -> list of per-class numpy arrays (rtransforms.py:10-25)."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 6), dtype=np.float32) for _ in range(num_classes)]
    bboxes = bboxes.cpu().numpy()
    labels = labels.cpu().numpy()
    return [bboxes[labels == i, :] for i in range(num_classes)]


def rbbox2roi(bbox_list):
    """list of (n,5+) boxes -> (sum n, 6) [batch_ind, cx, cy, w, h, a] (rtransforms.py:28-46)."""
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :5]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 6))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)
