"""utils.lua (hot subset) backed by the C ABI: nms, nms_dense, bbox_vote, convertFrom, keep_top_k."""
from __future__ import annotations

from typing import List

import numpy as np

from ._lib import Context


def nms(ctx: Context, boxes, overlap: float) -> np.ndarray:
    """utils.nms (utils.lua:29-33): N x 5 -> kept ROWS K x 5 in the order nms.c emits them."""
    b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 5)
    return b[ctx.nms(b, overlap)]


def nms_indices(ctx: Context, boxes, overlap: float) -> np.ndarray:
    return ctx.nms(boxes, overlap)


def nms_dense(ctx: Context, boxes, overlap: float) -> np.ndarray:
    """utils.nms_dense (utils.lua:402-462): 0-based indices (the reference returns 1-based LongTensor)."""
    return ctx.nms_dense(boxes, overlap)


def bbox_vote(ctx: Context, nms_boxes, scored_boxes, overlap: float) -> np.ndarray:
    return ctx.bbox_vote(nms_boxes, scored_boxes, overlap)


def convertFrom(ctx: Context, bbox, y) -> np.ndarray:
    """utils.convertFrom tensor branch (utils.lua:226-246); y may hold several class blocks of 4."""
    bbox = np.ascontiguousarray(bbox, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    if bbox.shape[0] != y.shape[0] or y.shape[1] % 4 != 0 or bbox.shape[1] != 4:
        raise ValueError("convertFrom: size mismatch")        # utils.lua:227-230
    return ctx.bbox_decode(y, bbox)


def keep_top_k(boxes: List[np.ndarray], top_k: int):
    """utils.keep_top_k (utils.lua:75-96): keep rows with score >= the top_k-th score over all classes."""
    nz = [b for b in boxes if b.size]
    if not nz:
        return boxes, 0
    scores = np.sort(np.concatenate([b[:, -1] for b in nz]))[::-1]
    thresh = scores[min(len(scores), top_k) - 1]
    return [b[b[:, -1] >= thresh] if b.size else b for b in boxes], float(thresh)
