"""TEST INFRASTRUCTURE ONLY — float32 numpy restatement of the torch box transforms in
r3det/core/bbox/rtransforms.py (poly2obb_v1/2/3 :190-277, obb2poly_v1/2/3 :367-440, obb2hbb_v1/2/3 :443-537,
hbb2obb_v1/2/3 :540-592, obb2xyxy_v1/2/3 :595-651, norm_angle :789-805).  Pinned against the reference
functions themselves through tests/golden/transforms_ref.npz.  Never imported by the product."""
import numpy as np

F = np.float32
PI = F(np.pi)
HPI = F(np.pi * 0.5)
QPI = F(np.pi / 4)


def _rem(a, b):
    """torch.remainder (sign of divisor) in float32."""
    a = np.asarray(a, F)
    m = np.fmod(a, F(b)).astype(F)
    fix = (m != 0) & ((F(b) < 0) != (m < 0))
    return np.where(fix, m + F(b), m).astype(F)


def norm_angle(a, v):
    if v == 'v2':
        return (_rem(a + QPI, PI) - QPI).astype(F)
    if v == 'v3':
        return (_rem(a + HPI, PI) - HPI).astype(F)
    return a


def obb2poly(b, v):
    b = np.asarray(b, F)
    x, y, w, h, a = (b[:, i] for i in range(5))
    c, s = np.cos(a).astype(F), np.sin(a).astype(F)
    if v == 'v1':
        wx, wy = w / F(2) * c, w / F(2) * s
        hx, hy = -h / F(2) * s, h / F(2) * c
        return np.stack([x - wx - hx, y - wy - hy, x + wx - hx, y + wy - hy,
                         x + wx + hx, y + wy + hy, x - wx + hx, y - wy + hy], -1).astype(F)
    tlx, tly, brx, bry = -w * F(0.5), -h * F(0.5), w * F(0.5), h * F(0.5)
    rx = np.stack([tlx, brx, brx, tlx], 1); ry = np.stack([tly, tly, bry, bry], 1)
    px = c[:, None] * rx - s[:, None] * ry + x[:, None]
    py = s[:, None] * rx + c[:, None] * ry + y[:, None]
    return np.stack([px, py], -1).reshape(len(b), 8).astype(F)


def poly2obb(p, v):
    p = np.asarray(p, F).reshape(-1, 8)
    if v == 'v1':
        cx = (p[:, 0] + p[:, 2] + p[:, 4] + p[:, 6]) / F(4)
        cy = (p[:, 1] + p[:, 3] + p[:, 5] + p[:, 7]) / F(4)
        _w = np.sqrt((p[:, 0] - p[:, 2]) ** 2 + (p[:, 1] - p[:, 3]) ** 2).astype(F)
        _h = np.sqrt((p[:, 2] - p[:, 4]) ** 2 + (p[:, 3] - p[:, 5]) ** 2).astype(F)
        th = np.arctan2(-(p[:, 2] - p[:, 0]), p[:, 3] - p[:, 1]).astype(F)
        odd = _rem(np.floor(th / (-HPI)), 2) == 0
        return np.stack([cx, cy, np.where(odd, _h, _w), np.where(odd, _w, _h), _rem(th, -HPI)], 1).astype(F)
    e1 = np.sqrt((p[:, 0] - p[:, 2]) ** 2 + (p[:, 1] - p[:, 3]) ** 2).astype(F)
    e2 = np.sqrt((p[:, 2] - p[:, 4]) ** 2 + (p[:, 3] - p[:, 5]) ** 2).astype(F)
    a1 = np.arctan2(p[:, 3] - p[:, 1], p[:, 2] - p[:, 0]).astype(F)
    a2 = np.arctan2(p[:, 7] - p[:, 1], p[:, 6] - p[:, 0]).astype(F)
    ang = norm_angle(np.where(e1 > e2, a1, a2).astype(F), v)
    return np.stack([(p[:, 0] + p[:, 4]) / F(2), (p[:, 1] + p[:, 5]) / F(2), np.maximum(e1, e2), np.minimum(e1, e2), ang], 1).astype(F)


def obb2xyxy(b, v):
    b = np.asarray(b, F)
    x, y, w, h, a = (b[:, i] for i in range(5))
    c, s = np.cos(a).astype(F), np.sin(a).astype(F)
    if v == 'v1':
        dw, dh = c * w - s * h, -s * w + c * h
        return np.stack([x - dw / F(2), y - dh / F(2), x + dw / F(2), y + dh / F(2)], -1).astype(F)
    if v == 'v2':
        p = obb2poly(b, 'v2')
        return np.stack([p[:, 0::2].min(1), p[:, 1::2].min(1), p[:, 0::2].max(1), p[:, 1::2].max(1)], 1).astype(F)
    xb = np.abs(w / F(2) * c) + np.abs(h / F(2) * s)
    yb = np.abs(w / F(2) * s) + np.abs(h / F(2) * c)
    return np.stack([x - xb, y - yb, x + xb, y + yb], -1).astype(F)


def obb2hbb(b, v):
    b = np.asarray(b, F)
    x, y, w, h, a = (b[:, i] for i in range(5))
    if v == 'v1':
        c, s = np.cos(a).astype(F), np.sin(a).astype(F)
        return np.stack([x, y, -s * w + c * h, c * w - s * h, np.full_like(x, -HPI)], 1).astype(F)
    if v == 'v2':
        bb = obb2xyxy(b, 'v2')
        xc, yc = (bb[:, 2] + bb[:, 0]) / F(2), (bb[:, 3] + bb[:, 1]) / F(2)
        e1, e2 = np.abs(bb[:, 2] - bb[:, 0]), np.abs(bb[:, 3] - bb[:, 1])
        sw = e1 < e2
        return np.stack([xc, yc, np.where(sw, e2, e1), np.where(sw, e1, e2), np.where(sw, HPI, F(0))], 1).astype(F)
    bb = obb2xyxy(b, 'v3')
    _x, _y = (bb[:, 0] + bb[:, 2]) * F(0.5), (bb[:, 1] + bb[:, 3]) * F(0.5)
    _w, _h = bb[:, 2] - bb[:, 0], bb[:, 3] - bb[:, 1]
    ok = _w >= _h
    return np.stack([_x, _y, np.where(ok, _w, _h), np.where(ok, _h, _w), np.where(ok, F(0), -HPI)], 1).astype(F)


def hbb2obb(hb, v):
    hb = np.asarray(hb, F)
    x, y = (hb[:, 0] + hb[:, 2]) * F(0.5), (hb[:, 1] + hb[:, 3]) * F(0.5)
    w, h = hb[:, 2] - hb[:, 0], hb[:, 3] - hb[:, 1]
    if v == 'v1':   # the reference slices with 0::4, so (N, 4) input yields (N, 1, 5) (rtransforms.py:548-554)
        return np.stack([x, y, h, w, np.full_like(x, -HPI)], 1).astype(F)[:, None, :]
    ok = w >= h
    alt = HPI if v == 'v2' else -HPI
    return np.stack([x, y, np.where(ok, w, h), np.where(ok, h, w), np.where(ok, F(0), alt)], 1).astype(F)
