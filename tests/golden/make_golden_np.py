"""Generate tests/golden/transforms_np_ref.npz from the reference's own numpy transforms (obb2poly_np, poly2obb_np,
get_best_begin_point in r3det/core/bbox/rtransforms.py), imported from /root/reference in the build container.
`np.float` (removed from numpy 1.24; used by poly2obb_np_v2, rtransforms.py:326-333) is aliased to float for the run.

    python tests/golden/make_golden_np.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)


def main():
    if not hasattr(np, "float"):
        np.float = float
    rt = mg.load_reference_python()["rtransforms"]
    rng = np.random.default_rng(20260301)
    out = {}
    for v in ("v1", "v2", "v3"):
        obb = np.concatenate([mg.rand_obb(120, rng, mg.AR[v], 3, 500), rng.uniform(0.05, 1, (120, 1)).astype(np.float32)], 1)
        out[f"{v}_obb"] = obb
        polys = rt.obb2poly_np(obb, v)
        out[f"{v}_obb2poly_np"] = np.asarray(polys)
        back = [rt.poly2obb_np(p[:8], v) for p in np.asarray(polys, np.float32)]
        out[f"{v}_poly2obb_np_ok"] = np.array([b is not None for b in back])
        out[f"{v}_poly2obb_np"] = np.array([b if b is not None else (0,) * 5 for b in back], np.float64)
        quad = np.asarray(polys, np.float32)[:, :8] + rng.normal(0, 2, (120, 8)).astype(np.float32)
        qb = [rt.poly2obb_np(p, v) for p in quad]
        out[f"{v}_quad"] = quad
        out[f"{v}_quad2obb_np_ok"] = np.array([b is not None for b in qb])
        out[f"{v}_quad2obb_np"] = np.array([b if b is not None else (0,) * 5 for b in qb], np.float64)
    np.savez_compressed(os.path.join(HERE, "transforms_np_ref.npz"), **out)
    print("transforms_np_ref.npz", os.path.getsize(os.path.join(HERE, "transforms_np_ref.npz")))


if __name__ == "__main__":
    main()
