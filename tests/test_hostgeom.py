"""The PRODUCT geometry headers (csrc/geom.cuh, csrc/emu.cuh) compiled for the host and swept against the
oracle: millions of pairs per regime without a GPU.  Gate (north_star): IoU within 1e-5 absolute."""
import numpy as np
import pytest

from tests.util import clustered, rand_obb
from oracle import port
from tests import hostgeom as hg

TOL = 1e-5


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_restatement_is_bitexact(v):
    a, b = rand_obb(500, 1, v), rand_obb(400, 2, v)
    assert np.array_equal(hg.emu_matrix(a, b, v), port.iou_matrix(a, b, v, wrapper_mask=False))
    c, _, _ = clustered(300, 3, v)
    assert np.array_equal(hg.emu_matrix(c, c, v, "iof"), port.iou_matrix(c, c, v, "iof", wrapper_mask=False))


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
@pytest.mark.parametrize("mode", ["iou", "iof"])
def test_fast_path_within_tolerance(v, mode):
    regimes = {
        "random": (rand_obb(3000, 11, v), rand_obb(1000, 12, v)),
        "small": (rand_obb(1500, 13, v, 0.5, 8, 64), rand_obb(1000, 14, v, 0.5, 8, 64)),
        "aspect": (rand_obb(1500, 15, v, 2, 1000), rand_obb(800, 16, v, 2, 1000)),
    }
    c, _, _ = clustered(2000, 17, v)
    regimes["clustered"] = (c, c)
    d = c[:800].copy(); e = d.copy(); e[:, :2] += np.random.default_rng(0).normal(0, 1e-3, (800, 2)).astype(np.float32)
    regimes["near_identical"] = (d, e)
    an = rand_obb(1500, 18, v, 16, 400); an[:, 4] = 0
    gt = rand_obb(500, 19, v, 10, 300); gt[:, 4] = np.float32(-np.pi / 2)
    regimes["axis_aligned"] = (gt, an)
    for name, (a, b) in regimes.items():
        st = {}
        fast = hg.iou_matrix(a, b, v, mode, stats=st)
        want = port.iou_matrix(a, b, v, mode, wrapper_mask=False)
        err = np.abs(fast - want).max()
        assert err <= TOL, f"{v}/{mode}/{name}: max |fast - oracle| = {err:.3g}"
        # the separating-axis filter must never zero a pair the reference scores above noise
        assert want[fast == 0].max(initial=0) < 1e-6


def test_fast_path_matches_exact_geometry():
    """Pure geometry (no strict fallback, no v1 corner re-fit): the clamped-boundary integral against the
    float64 Sutherland-Hodgman ground truth."""
    for v in ("v1", "v3"):
        a, b = rand_obb(3000, 21, v), rand_obb(1000, 22, v)
        ex = port.iou_exact(a, b, v)
        fast = hg.iou_matrix(a, b, v, tau=0.0, refit=False)
        assert np.abs(fast - ex).max() < 2e-6
