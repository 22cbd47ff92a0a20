"""CPU: host-side numpy helpers of the offline submission path (obb2poly_np / poly2obb_np / patch bookkeeping / Task-1
writer) against vectors produced by the reference's own functions (tests/golden/make_golden_np.py)."""
import os
import zipfile

import numpy as np
import pytest

from tests.util import golden


def _mod():
    import r3det_b200
    return r3det_b200


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_np_transforms_golden(v):
    R = _mod()
    g = golden("transforms_np_ref.npz")
    got = R.obb2poly_np(g[f"{v}_obb"], v)
    assert got.shape == g[f"{v}_obb2poly_np"].shape
    assert np.abs(got - g[f"{v}_obb2poly_np"]).max() < 2e-4
    pytest.importorskip("cv2")
    for key_in, key_out in ((f"{v}_obb2poly_np", f"{v}_poly2obb_np"), (f"{v}_quad", f"{v}_quad2obb_np")):
        src = np.asarray(g[key_in], np.float32)[:, :8]
        for p, ok, want in zip(src, g[key_out + "_ok"], g[key_out]):
            r = R.poly2obb_np(p, v)
            assert (r is not None) == bool(ok)
            if ok:
                assert np.allclose(np.array(r, np.float64), want, rtol=0, atol=1e-4)
    with pytest.raises(NotImplementedError):
        R.obb2poly_np(g[f"{v}_obb"], "v9")


def test_patch_bookkeeping_and_writer(tmp_path):
    R = _mod()
    ds = R.dota_submission
    assert ds.patch_origin("P0006__1__0___824") == ("P0006", 0, 824)
    assert ds.patch_origin("P2__0.5__1024___0") == ("P2", 1024, 0)
    res = [[np.array([[10, 20, 30, 8, -0.3, 0.9]], np.float32), np.zeros((0, 6), np.float32)],
           [np.zeros((0, 6), np.float32), np.array([[5, 5, 12, 6, -1.0, 0.5], [7, 7, 9, 4, -0.2, 0.4]], np.float32)]]
    col = ds.collect_patches(res, ["A__1__0___824", "A__1__512___0"])
    assert list(col) == ["A"] and col["A"].shape == (3, 7)
    assert np.allclose(col["A"][0], [0, 10, 844, 30, 8, -0.3, 0.9]) and np.allclose(col["A"][1], [1, 517, 5, 12, 6, -1.0, 0.5])
    out = str(tmp_path / "sub")
    per_cls = [col["A"][col["A"][:, 0] == c][:, 1:] for c in range(2)]
    files = ds.write_task1(out, ["A"], [per_cls], ["plane", "ship"], "v1")
    assert [os.path.basename(f) for f in files] == ["Task1_plane.txt", "Task1_ship.txt"]
    lines = open(files[1]).read().strip().split("\n")
    assert len(lines) == 2 and lines[0].split()[0] == "A" and len(lines[0].split()) == 10
    want = R.obb2poly_np(per_cls[1], "v1")[0]
    assert lines[0].split()[1] == str(want[-1]) and lines[0].split()[2:] == [f"{p:.2f}" for p in want[:-1]]
    assert sorted(zipfile.ZipFile(os.path.join(out, "sub.zip")).namelist()) == ["Task1_plane.txt", "Task1_ship.txt"]
    with pytest.raises(ValueError):
        ds.write_task1(out, [], [], ["plane"], "v1")
