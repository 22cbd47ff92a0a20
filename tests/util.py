"""Seeded synthetic inputs shared by the tests (SURVEY.md §8d generators)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
AR = {'v1': (-np.pi / 2, 0), 'v2': (-np.pi / 4, 3 * np.pi / 4), 'v3': (-np.pi / 2, np.pi / 2)}


def rand_obb(n, seed, version='v1', lo=8, hi=512, span=1024):
    """SURVEY.md §8d `rand_obb`: centres U(0,span), log-uniform sizes, angle uniform over the version's range."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, span, n); cy = rng.uniform(0, span, n)
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n)); h = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    a = rng.uniform(*AR[version], n)
    return np.stack([cx, cy, w, h, a], 1).astype(np.float32)


def clustered(K, seed, version='v1', ncls=15):
    """SURVEY.md §8d clustered NMS candidates: K/10 seed boxes + jitter, label = seed mod ncls, distinct scores."""
    rng = np.random.default_rng(seed)
    seeds = rand_obb(max(K // 10, 1), seed + 1000, version, 12, 200)
    idx = rng.integers(0, len(seeds), K)
    b = seeds[idx].copy()
    b[:, 0:2] += rng.normal(0, 4, (K, 2)); b[:, 4] += rng.normal(0, 0.05, K)
    b[:, 2:4] *= np.exp(rng.normal(0, 0.1, (K, 2)))
    labels = (idx % ncls).astype(np.int64)
    scores = rng.permutation(np.linspace(0.05, 1, K)).astype(np.float32)
    return b.astype(np.float32), scores, labels


def anchors_1024():
    """RAnchorGenerator semantics (r3det/core/anchor/ranchor_generator.py:30-39 on mmdet's AnchorGenerator):
    strides 8..128, octave_base_scale 4, 3 scales/octave, ratios [1, .5, 2], centres (i*s, j*s), theta 0."""
    out = []
    for s in (8, 16, 32, 64, 128):
        n = 1024 // s
        scales = np.array([4 * 2 ** (k / 3) for k in range(3)])
        ratios = np.array([1.0, 0.5, 2.0])
        h_r, w_r = np.sqrt(ratios), 1 / np.sqrt(ratios)
        ws = (s * w_r[:, None] * scales[None, :]).reshape(-1)
        hs = (s * h_r[:, None] * scales[None, :]).reshape(-1)
        ys, xs = np.meshgrid(np.arange(n) * s, np.arange(n) * s, indexing='ij')
        ctr = np.stack([xs.reshape(-1), ys.reshape(-1)], 1)
        a = np.zeros((n * n, 9, 5), np.float32)
        a[:, :, :2] = ctr[:, None, :]
        a[:, :, 2] = ws[None]; a[:, :, 3] = hs[None]
        out.append(a.reshape(-1, 5))
    return np.concatenate(out).astype(np.float32)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


