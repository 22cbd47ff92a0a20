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
    L.require_no_grad('rtransforms.' + fn_name.replace('r3g_', '').replace('_f32', ''), x)
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


# ------------------------------------------------------------------------------------------------------------------
# numpy variants used on the host side of the offline DOTA submission (rtransforms.py:70-87, 130-147, 280-364,
# 654-786).  Host code in the reference, host code here: vectorised numpy instead of per-box Python loops.
def _best_begin_point(polys9):
    """Rotate each polygon's vertex order so that it starts nearest the top-left of its bounding box
    (get_best_begin_point(_single), rtransforms.py:742-786): float64 in, float64 out, first minimum wins."""
    p = np.asarray(polys9, np.float64)
    if p.shape[0] == 0:
        return p.reshape(0, 9)
    pts = p[:, :8].reshape(-1, 4, 2)
    lo, hi = pts.min(1), pts.max(1)
    dst = np.stack([lo, np.stack([hi[:, 0], lo[:, 1]], -1), hi, np.stack([lo[:, 0], hi[:, 1]], -1)], 1)    # (n, 4, 2)
    cost = np.stack([np.sqrt(((np.roll(pts, -s, axis=1) - dst) ** 2).sum(-1)).sum(-1) for s in range(4)], 1)
    start = cost.argmin(1)
    idx = (start[:, None] + np.arange(4)[None]) % 4
    out = np.take_along_axis(pts, idx[:, :, None], axis=1).reshape(-1, 8)
    return np.concatenate([out, p[:, 8:9]], 1)


def obb2poly_np(rbboxes, version='v1'):
    """(n, 6) [x_ctr, y_ctr, w, h, angle, score] -> (n, 9) [x0, y0, ..., x3, y3, score] (rtransforms.py:130-147)."""
    r = np.asarray(rbboxes)
    if version == 'v1':                                                                    # :654-676
        x, y, w, h, a, score = (r[:, i] for i in range(6))
        c, s = np.cos(a), np.sin(a)
        wx, wy, hx, hy = w / 2 * c, w / 2 * s, -h / 2 * s, h / 2 * c
        return np.stack([x - wx - hx, y - wy - hy, x + wx - hx, y + wy - hy,
                         x + wx + hx, y + wy + hy, x - wx + hx, y - wy + hy, score], axis=-1)
    if version == 'v2':                                                                    # :679-702
        if r.shape[0] == 0:
            return np.zeros((0,), np.float64)        # np.array([]) of the reference's loop
        r32 = r.astype(np.float32, copy=False)
        x, y, w, h, a, score = (r32[:, i] for i in range(6))
        c, s = np.cos(a), np.sin(a)
        rx = np.stack([-w / 2, w / 2, w / 2, -w / 2], 1); ry = np.stack([-h / 2, -h / 2, h / 2, h / 2], 1)
        px = c[:, None] * rx - s[:, None] * ry + x[:, None]
        py = s[:, None] * rx + c[:, None] * ry + y[:, None]
        polys = np.stack([px[:, 0], py[:, 0], px[:, 1], py[:, 1], px[:, 2], py[:, 2], px[:, 3], py[:, 3], score], 1).astype(np.float32)
        return _best_begin_point(polys)
    if version == 'v3':                                                                    # :705-725
        try:
            center, w, h, theta, score = np.split(r, (2, 3, 4, 5), axis=-1)
        except Exception:  # noqa: BLE001  (the reference answers malformed input with one zero row)
            return np.zeros((1, 9))
        c, s = np.cos(theta), np.sin(theta)
        v1 = np.concatenate([w / 2 * c, -w / 2 * s], axis=-1)
        v2 = np.concatenate([-h / 2 * s, -h / 2 * c], axis=-1)
        return np.concatenate([center + v1 + v2, center + v1 - v2, center - v1 - v2, center - v1 + v2, score], axis=-1)
    raise NotImplementedError


def poly2obb_np(polys, version='v1'):
    """ONE polygon [x0, y0, ..., x3, y3] -> (x_ctr, y_ctr, w, h, angle), or None for boxes thinner than 2 px
    (rtransforms.py:70-87, 280-364).  v1 / v3 go through cv2.minAreaRect like the reference."""
    if version == 'v2':                                                                    # :306-337
        q = np.array(polys[:8], dtype=np.float32)
        e1 = np.sqrt((q[0] - q[2]) * (q[0] - q[2]) + (q[1] - q[3]) * (q[1] - q[3]))
        e2 = np.sqrt((q[2] - q[4]) * (q[2] - q[4]) + (q[3] - q[5]) * (q[3] - q[5]))
        if e1 < 2 or e2 < 2:
            return None
        if e1 > e2:
            ang = np.arctan2(float(q[3] - q[1]), float(q[2] - q[0]))
        else:
            ang = np.arctan2(float(q[7] - q[1]), float(q[6] - q[0]))
        return float(q[0] + q[4]) / 2, float(q[1] + q[5]) / 2, max(e1, e2), min(e1, e2), norm_angle(ang, 'v2')
    if version not in ('v1', 'v3'):
        raise NotImplementedError
    import cv2
    (x, y), (w, h), a = cv2.minAreaRect(np.array(polys).reshape((4, 2)))
    if w < 2 or h < 2:
        return None
    if version == 'v1':                                                                    # :280-303
        while not 0 > a >= -90:
            a, w, h = (a - 90, h, w) if a >= 0 else (a + 90, h, w)
        a = a / 180 * np.pi
        assert 0 > a >= -np.pi / 2
        return x, y, w, h, a
    a = -a / 180 * np.pi                                                                   # :340-364
    if w < h:
        w, h = h, w
        a += np.pi / 2
    while not np.pi / 2 > a >= -np.pi / 2:
        a = a - np.pi if a >= np.pi / 2 else a + np.pi
    return x, y, w, h, a
