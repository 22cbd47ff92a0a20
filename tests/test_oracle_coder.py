"""CPU: the numpy restatement of the coder / get_bboxes tail (oracle/coder_np.py) against vectors produced by the
reference's own Python (tests/golden/coder_ref.npz, generator: tests/golden/make_golden_coder.py)."""
import numpy as np
import pytest

from oracle import coder_np as cn
from tests.util import golden

SIZES = [(16, 20, 8), (8, 10, 16), (4, 5, 32)]
ZERO, ONE = (0.0,) * 5, (1.0,) * 5


def close(a, b, tol=2e-5):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= tol, err.max()


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_encode_decode(v):
    g = golden("coder_ref.npz")
    mb, sb = tuple(g["means_b"].tolist()), tuple(g["stds_b"].tolist())
    for tag, (m, s) in (("a", (ZERO, ONE)), ("b", (mb, sb))):
        close(cn.bbox2delta(g[f"{v}_prop"], g[f"{v}_gt"], m, s, v), g[f"{v}_{tag}_encode"], 5e-5)
        close(cn.delta2bbox(g[f"{v}_prop"], g[f"{v}_deltas"], m, s, v), g[f"{v}_{tag}_decode"])
        close(cn.delta2bbox(g[f"{v}_prop"], g[f"{v}_deltas3"], m, s, v), g[f"{v}_{tag}_decode3"])
    close(cn.delta2bbox(g[f"{v}_prop"], g[f"{v}_deltas"], mb, sb, v, max_shape=(512, 640, 3)), g[f"{v}_decode_clamped"])
    close(cn.delta2bbox(g[f"{v}_prop"], g[f"{v}_deltas"], mb, sb, v, add_ctr_clamp=True, ctr_clamp=8), g[f"{v}_decode_ctr"])


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
@pytest.mark.parametrize("A", [1, 3])
def test_heads(v, A):
    g = golden("coder_ref.npz")
    tag = f"{v}_A{A}"
    coder = dict(means=ZERO, stds=tuple(g[f"{tag}_stds"].tolist()), variant=v)
    cls = [g[f"{tag}_cls{i}"] for i in range(3)]; reg = [g[f"{tag}_reg{i}"] for i in range(3)]; anc = [g[f"{tag}_anc{i}"] for i in range(3)]
    for r in (0, 1):
        b, s = cn.select_decode(cls, reg, anc, (140, 170, 3), [1.25, 1.5, 1.25, 1.5], 100, 15, coder, rescale=bool(r))
        close(b, g[f"{tag}_r{r}_mlvl_bboxes"]); close(s, g[f"{tag}_r{r}_mlvl_scores"], 1e-6)
    for img in range(2):
        for i in range(3):
            c_ = cls[i] if img == 0 else -cls[i][::-1]
            r_ = reg[i] if img == 0 else -reg[i]
            f = cn.filter_bboxes(c_, r_, anc[i], A, 15, coder)
            close(f, g[f"{tag}_filter_img{img}_lvl{i}"])
            if A == 1:
                close(cn.refine_bboxes(r_, g[f"{tag}_filter_img{img}_lvl{i}"], coder), g[f"{tag}_refine_img{img}_lvl{i}"])
