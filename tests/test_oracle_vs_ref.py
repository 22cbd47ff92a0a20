"""The oracle against the unmodified reference binaries (oracle/_ref, built from /root/reference in place) on
seeded sweeps.  Skipped when the binaries are absent; the committed golden vectors cover the same ground."""
import numpy as np
import pytest

from tests.util import clustered, rand_obb
from oracle import port, ref

pytestmark = pytest.mark.skipif(not ref.available("libref_v1.so"), reason="oracle/_ref not built")


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_iou_sweep_bitexact(v):
    for seed, (lo, hi, span) in enumerate([(8, 512, 1024), (0.5, 8, 64), (2, 1000, 1024)]):
        a, b = rand_obb(600, seed, v, lo, hi, span), rand_obb(400, seed + 50, v, lo, hi, span)
        got = port.iou_matrix(a, b, v, wrapper_mask=False, sort_kind=1)
        want = {"v1": lambda: ref.v1_iou(a, b), "v2": lambda: ref.v2_iou(a, b), "v3": lambda: ref.v3_iou(a, b)}[v]()
        assert np.array_equal(got, want)


def test_nms_sweep():
    for v in ("v1", "v2", "v3"):
        b, s, l = clustered(1500, 7, v)
        if v == "v1":
            assert np.array_equal(np.sort(port.nms(b, s, 0.1, v, inclusive=True)), ref.v1_nms(np.concatenate([b, s[:, None]], 1), 0.1))
        elif v == "v3":
            assert np.array_equal(port.nms(b, s, 0.1, v, inclusive=True, sort_kind=1), ref.v3_nms(b, s, 0.1))
        else:
            assert np.array_equal(port.nms(b, s, 0.1, v, labels=l.astype(np.float32), inclusive=True, sort_kind=1), ref.v2_nms(b, s, l, 0.1))
