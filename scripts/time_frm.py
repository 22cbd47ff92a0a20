"""GPU timing probe for FRM (batch 8, 256 channels, five FPN levels of a 1024^2 patch): forward, backward (plan + apply) and
apply alone, points 1 / 5, plus the per-level forward.  R3G_FRM_TMA=0 selects the gather kernels.  Prints one JSON line."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import r3det_b200 as R
from r3det_b200.fr import FrmBackwardPlan, frm_backward_multi, frm_forward, frm_forward_multi
dev = torch.device("cuda:0")
rng = np.random.default_rng(4)
xs, bts, scales = [], [], []
for H, stride in ((128, 8), (64, 16), (32, 32), (16, 64), (8, 128)):
    xs.append(torch.randn((8, 256, H, H), device=dev))
    ys_, xs_ = np.meshgrid(np.arange(H) * stride, np.arange(H) * stride, indexing="ij")
    ctr = np.stack([xs_, ys_], -1).reshape(-1, 2).astype(np.float32)
    bx = np.zeros((8, H * H, 5), np.float32)
    bx[:, :, :2] = ctr[None] + rng.normal(0, stride, (8, H * H, 2))
    bx[:, :, 2:4] = np.exp(rng.uniform(np.log(stride), np.log(8 * stride), (8, H * H, 2)))
    bx[:, :, 4] = rng.uniform(-np.pi / 2, 0, (8, H * H))
    bts.append(torch.from_numpy(bx.reshape(-1, 5)).to(dev)); scales.append(1.0 / stride)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / iters


elems = sum(x.numel() for x in xs)
res = {"tma": os.environ.get("R3G_FRM_TMA", "1"), "elements": elems}
for P in (1, 5):
    tf = timeit(lambda: frm_forward_multi(xs, bts, scales, P))
    tb = timeit(lambda: frm_backward_multi(xs, bts, scales, P))
    plan = FrmBackwardPlan([tuple(x.shape) for x in xs], bts, scales, P)
    torch.cuda.synchronize()
    ta = timeit(lambda: plan.apply(xs))
    t0 = timeit(lambda: frm_forward(xs[0], bts[0], scales[0], P))
    o = frm_forward_multi(xs, bts, scales, P)
    res[f"P{P}"] = {"fwd_ms": tf, "fwd_frac": elems * 8 / tf / 1e6 / 6453.7, "bwd_ms": tb, "apply_ms": ta,
                    "apply_frac": elems * 8 / ta / 1e6 / 6453.7, "fwd_level0_ms": t0, "checksum": float(sum(t.double().sum() for t in o))}
print(json.dumps(res))
