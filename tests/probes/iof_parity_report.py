"""GPU box: the v1 IoF parity footnote, pair by pair.  tests/test_iou_gpu.py gates v1 IoF at 2e-5 (not 1e-5) on the goldens that
contain sub-8-px boxes.  This report lists every pair of tests/golden/iou_refcuda.npz (the reference's own CUDA kernel, run on a
B200) where |ours - reference CUDA| > 1e-5, next to the disagreement of the REFERENCE WITH ITSELF on the same pair: its CPU build
(rnms_cpu.cpp host templates, oracle/_ref/libref_v1.so) against its CUDA build.  Output: profiles/r02_v1_iof_parity.txt"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import r3det_b200 as R
from oracle import ref
from tests.util import golden

dev = torch.device("cuda:0")
from oracle import refcuda
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
out = []
g = golden("iou_refcuda.npz")
g2 = golden("iou_ref.npz")
for gname, b1, b2, name, iof, cuda in (("iou_refcuda.npz", g["b1"], g["b2"], "IoU", False, g["v1_iou"]),
                                       ("iou_refcuda.npz", g["b1"], g["b2"], "IoF", True, g["v1_iof"]),
                                       ("iou_ref.npz (sub-8-px boxes)", g2["v1_b1"], g2["v1_b2"], "IoU", False, None),
                                       ("iou_ref.npz (sub-8-px boxes)", g2["v1_b1"], g2["v1_b2"], "IoF", True, None)):
    ours = R.rbbox_iou(t(b1), t(b2), False, iof).cpu().numpy()
    if cuda is None:        # this fixture holds the reference's CPU output: run its CUDA kernel here
        o = torch.empty((b1.shape[0], b2.shape[0]), device=dev)
        refcuda.v1_iou_ms(t(b1), t(b2), o, 1, iof)
        cuda = o.cpu().numpy()
    cpu = ref.v1_iou(b1, b2, iof=iof)
    out.append(f"#### fixture {gname}")
    d_ours, d_self = np.abs(ours - cuda), np.abs(cpu - cuda)
    ov = cuda > 0
    out.append(f"== v1 {name}: {b1.shape[0]} x {b2.shape[0]} boxes, {int(ov.sum())} overlapping pairs (reference CUDA kernel = golden)")
    out.append(f"   ours vs reference CUDA : max {d_ours.max():.3e}   pairs > 1e-5: {int((d_ours > 1e-5).sum())}   > 2e-5: {int((d_ours > 2e-5).sum())}")
    out.append(f"   reference CPU vs CUDA  : max {d_self.max():.3e}   pairs > 1e-5: {int((d_self > 1e-5).sum())}   > 2e-5: {int((d_self > 2e-5).sum())}")
    idx = np.argwhere(d_ours > 1e-5)
    for i, j in idx[:40]:
        out.append(f"   pair ({i:3d},{j:3d}) first box w,h = {b1[i,2]:7.3f},{b1[i,3]:7.3f}  second w,h = {b2[j,2]:7.3f},{b2[j,3]:7.3f}  "
                   f"ref CUDA {cuda[i,j]:.7f}  ours {ours[i,j]:.7f} (diff {d_ours[i,j]:.2e})  ref CPU {cpu[i,j]:.7f} (diff {d_self[i,j]:.2e})")
    if len(idx) > 40:
        out.append(f"   ... {len(idx) - 40} more")
    if len(idx):
        sizes = np.minimum(b1[idx[:, 0], 2], b1[idx[:, 0], 3])
        out.append(f"   every such pair has a first box (the IoF denominator) with min(w, h) <= {sizes.max():.2f} px")
txt = "\n".join(out)
print(txt)
os.makedirs(os.path.join(ROOT, "gpurun_out", "r2"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "r2", "r02_v1_iof_parity.txt"), "w").write(txt + "\n")
