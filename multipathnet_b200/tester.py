"""fbcoco.Tester_FRCNN:testOne mirror (Tester_FRCNN.lua:54-139) over the C ABI."""
from __future__ import annotations

from typing import List

import numpy as np

from ._lib import Model
from .image_detect import ImageDetect
from . import utils as U


class Tester:
    def __init__(self, model: Model, transformer, scale=None, max_size=None, nms_thresh: float = 0.3,
                 bbox_vote_thresh: float = 0.5, score_thresh: float = -1.5, bbox_voting: bool = False):
        self.detec = ImageDetect(model, transformer, scale, max_size)
        self.model = model
        self.nms_thresh = nms_thresh                 # opt.test_nms_threshold (Tester_FRCNN.lua:28)
        self.bbox_vote_thresh = bbox_vote_thresh     # the reference reads an unset field here; intended 0.5
        self.thresh = score_thresh                   # Tester_FRCNN.lua:50
        self.bbox_voting = bbox_voting
        self.num_classes = model.C - 1

    def testOne(self, im, boxes) -> List[np.ndarray]:
        """-> img_boxes: list over foreground classes of K_j x 5 [x1,y1,x2,y2,score] after NMS.
        One library call: trunk, heads, decode, clamp, per-class gather and batched NMS all on the GPU."""
        img, im_scale = self.detec.getImages(im)
        H0, W0 = im.shape[1], im.shape[2]
        scores, bboxes, keeps = self.model.detect_nms(img, boxes, im_scale, W0, H0, self.thresh, self.nms_thresh)
        out = []
        for j, k in enumerate(keeps, start=1):
            sb = np.concatenate([bboxes[k, 4 * j:4 * j + 4], scores[k, j:j + 1]], axis=1).astype(np.float32)
            if self.bbox_voting and len(k):
                sel = scores[:, j] > self.thresh
                allsb = np.concatenate([bboxes[sel, 4 * j:4 * j + 4], scores[sel, j:j + 1]], axis=1).astype(np.float32)
                sb = U.bbox_vote(self.model.ctx, sb, allsb, self.bbox_vote_thresh)
            out.append(sb)
        self.raw = (scores, bboxes)
        return out
