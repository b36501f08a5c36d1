"""Seeded synthetic workloads for the five BASELINE.json configs.

No dataset or proposal file is reachable (no network, data/ absent from the reference tree),
so images are uniform noise pushed through the model's transformer and proposals are random
boxes of the stated shape distribution. Everything is a pure function of the seed.
"""
from __future__ import annotations

import numpy as np

ROSS_MEAN = (102.9801, 115.9465, 122.7717)             # model_utils.lua:138-140
IMAGENET_MEAN = (0.48462227599918, 0.45624044862054, 0.40588363755159)   # model_utils.lua:143-155
IMAGENET_STD = (0.22889466674951, 0.22446679341259, 0.22495548344775)


def raw_image(h: int, w: int, seed: int) -> np.ndarray:
    """3 x H x W float in [0,1], RGB — what loaders/loader.lua:79 hands to detect()."""
    return np.random.default_rng(seed).random((3, h, w), dtype=np.float32)


def transform(im: np.ndarray, kind: str) -> np.ndarray:
    """fbcoco.ImageTransformer:updateOutput (modules/ImageTransformer.lua:19-33)."""
    if kind == "ross":                                  # swap {3,2,1}, x255, minus mean
        out = im[[2, 1, 0]].astype(np.float32) * np.float32(255.0)
        for c in range(3):
            out[c] -= np.float32(ROSS_MEAN[c])
        return out
    out = im.astype(np.float32).copy()
    for c in range(3):
        out[c] = (out[c] - np.float32(IMAGENET_MEAN[c])) / np.float32(IMAGENET_STD[c])
    return out


def random_boxes(n: int, img_h: int, img_w: int, seed: int, wmin=16, wmax=None, hmin=16, hmax=None) -> np.ndarray:
    """cfg 2: w ~ U[16, 0.8W], h ~ U[16, 0.8H], top-left uniform s.t. the box stays inside; 1-based [x1,y1,x2,y2]."""
    rng = np.random.default_rng(seed)
    wmax = wmax or 0.8 * img_w
    hmax = hmax or 0.8 * img_h
    w = rng.uniform(wmin, wmax, n)
    h = rng.uniform(hmin, hmax, n)
    x1 = 1 + rng.uniform(0, 1, n) * (img_w - w - 1)
    y1 = 1 + rng.uniform(0, 1, n) * (img_h - h - 1)
    return np.stack([x1, y1, x1 + w, y1 + h], 1).astype(np.float32)


def sharpmask_boxes(n: int, img_h: int, img_w: int, seed: int) -> np.ndarray:
    """cfg 3/4 'SharpMask-shaped' proposals (builder's definition): longest side 128/2^s,
    s in {-2.5..0.5 step .5} (demo.lua:23-25,49-50) x U[.75,1.25], aspect exp(U[ln 1/3, ln 3]), clipped."""
    rng = np.random.default_rng(seed)
    s = rng.choice(np.arange(-2.5, 0.51, 0.5), n)
    L = 128.0 / (2.0 ** s) * rng.uniform(0.75, 1.25, n)
    ar = np.exp(rng.uniform(np.log(1 / 3), np.log(3), n))
    w = np.where(ar >= 1, L, L * ar)
    h = np.where(ar >= 1, L / ar, L)
    cx = rng.uniform(1, img_w, n)
    cy = rng.uniform(1, img_h, n)
    x1 = np.clip(cx - w / 2, 1, img_w - 2); x2 = np.clip(cx + w / 2, x1 + 1, img_w)
    y1 = np.clip(cy - h / 2, 1, img_h - 2); y2 = np.clip(cy + h / 2, y1 + 1, img_h)
    return np.stack([x1, y1, x2, y2], 1).astype(np.float32)


def nms_sweep_boxes(n: int, ncls: int, seed: int, img_h=600, img_w=800, ties=False) -> np.ndarray:
    """cfg 5: ncls x n x 5 scored boxes; distinct scores unless ties=True (scores rounded to 1/20)."""
    rng = np.random.default_rng(seed)
    out = np.empty((ncls, n, 5), np.float32)
    for c in range(ncls):
        out[c, :, :4] = random_boxes(n, img_h, img_w, seed * 1000 + c, wmax=0.4 * img_w, hmax=0.4 * img_h)
        if ties:
            out[c, :, 4] = np.round(rng.random(n) * 20) / 20
        else:
            sc = rng.permutation(n).astype(np.float64) + rng.random(n) * 0.5      # distinct by construction
            out[c, :, 4] = (sc / n).astype(np.float32)
            assert len(np.unique(out[c, :, 4])) == n
    return out
