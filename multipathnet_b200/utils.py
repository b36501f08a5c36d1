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


def joinTable(tensors: List[np.ndarray], dim: int = 0) -> np.ndarray:
    """utils.joinTable (utils.lua:42-71): concatenate, skipping empty tensors."""
    nz = [np.asarray(t) for t in tensors if np.asarray(t).size]
    if not nz:
        return np.zeros((0,), np.float32)
    return np.concatenate(nz, axis=dim)


def transposeBoxes(aboxes_t: List[List[np.ndarray]], num_classes: int) -> List[List[np.ndarray]]:
    """Tester:transposeBoxes (Tester_FRCNN.lua:176-187): [image][class] -> [class][image]."""
    return [[aboxes_t[i][j] for i in range(len(aboxes_t))] for j in range(num_classes)]


def coco_results(aboxes: List[List[np.ndarray]], image_ids: List[int], category_ids: List[int]) -> np.ndarray:
    """testCoco.evaluate's result tensor (testCoco/init.lua:65-86): rows [image_id, x1-1, y1-1, w, h, score, category_id]
    with w = x2 - x1 and h = y2 - y1 (no +1, as the reference), classes outermost, images in order."""
    rows = []
    for j, per_img in enumerate(aboxes):
        for i, t in enumerate(per_img):
            t = np.asarray(t, np.float32)
            if t.ndim == 2 and t.shape[0] > 0:
                r = np.empty((t.shape[0], 7), np.float32)
                r[:, 0] = image_ids[i]; r[:, 1] = t[:, 0] - 1; r[:, 2] = t[:, 1] - 1
                r[:, 3] = t[:, 2] - t[:, 0]; r[:, 4] = t[:, 3] - t[:, 1]; r[:, 5] = t[:, 4]; r[:, 6] = category_ids[j]
                rows.append(r)
    return np.concatenate(rows, 0) if rows else np.zeros((0, 7), np.float32)
