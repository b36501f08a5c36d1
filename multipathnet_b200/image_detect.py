"""fbcoco.ImageDetect mirror (ImageDetect.lua) over the C ABI."""
from __future__ import annotations

import numpy as np

from ._lib import Model
from .modules import ImageTransformer


def _bilinear_resize(im: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """image.scale(im, w, h) bilinear (third-party `image` package, unpinned; SURVEY 8c "parity unpinned").
    Pixel-centre aligned sampling; identity when the size is unchanged (all benchmark configs)."""
    c, h, w = im.shape
    if (h, w) == (out_h, out_w):
        return im
    ys = (np.arange(out_h, dtype=np.float64) + 0.5) * h / out_h - 0.5
    xs = (np.arange(out_w, dtype=np.float64) + 0.5) * w / out_w - 0.5
    y0 = np.clip(np.floor(ys), 0, h - 1).astype(int); y1 = np.clip(y0 + 1, 0, h - 1)
    x0 = np.clip(np.floor(xs), 0, w - 1).astype(int); x1 = np.clip(x0 + 1, 0, w - 1)
    wy = np.clip(ys - y0, 0, 1)[None, :, None]; wx = np.clip(xs - x0, 0, 1)[None, None, :]
    a = im[:, y0][:, :, x0]; b = im[:, y0][:, :, x1]; c_ = im[:, y1][:, :, x0]; d = im[:, y1][:, :, x1]
    return ((a * (1 - wx) + b * wx) * (1 - wy) + (c_ * (1 - wx) + d * wx) * wy).astype(np.float32)


class ImageDetect:
    def __init__(self, model: Model, transformer: ImageTransformer, scale=None, max_size=None):
        if model is None:
            raise ValueError("must provide model!")           # ImageDetect.lua:13
        if transformer is None:
            raise ValueError("must provide transformer!")     # ImageDetect.lua:14
        self.model = model
        self.image_transformer = transformer
        self.scale = list(scale) if scale else [600]
        self.max_size = max_size or 1000
        if len(self.scale) != 1:
            # project_im_rois' multi-scale branch never fills rois (ImageDetect.lua:57-65): single scale only
            raise ValueError("only a single test scale is functional in the reference")

    def getImages(self, im):
        """ImageDetect.lua:22-52: transformer, scale to self.scale (capped by max_size)."""
        im = self.image_transformer.forward(im)
        h, w = im.shape[1], im.shape[2]
        smin, smax = min(h, w), max(h, w)
        im_scale = self.scale[0] / smin
        if round(im_scale * smax) > self.max_size:
            im_scale = self.max_size / smax
        # image.scale(im, w, h) receives float sizes and truncates them (Lua -> C long)
        out_h, out_w = int(h * im_scale + 1e-9), int(w * im_scale + 1e-9)
        return _bilinear_resize(im, out_h, out_w), float(im_scale)

    def detect(self, im, boxes, min_images=None, recompute_features=True):
        """-> (scores R x C float32, bboxes R x 4C float32) in original-image coordinates.
        `min_images` (DataParallelTable width) only replicated the image in the reference; ignored."""
        boxes = np.ascontiguousarray(boxes, np.float32)
        if boxes.ndim != 2 or boxes.shape[1] != 4:
            raise ValueError("boxes must be R x 4 [x1,y1,x2,y2]")
        if recompute_features:
            img, im_scale = self.getImages(im)
            self._im_scale = im_scale
        else:
            img, im_scale = None, self._im_scale
        return self.model.detect(img, boxes, im_scale, recompute_features)

    def computeRawOutputs(self, im, boxes, min_images=None, recompute_features=True):
        """ImageDetect.lua:137-153: model:forward on the projected ROIs (no decode / softmax)."""
        img, im_scale = self.getImages(im)
        rois = np.empty((boxes.shape[0], 5), np.float32)
        rois[:, 0] = 1
        rois[:, 1:] = (np.asarray(boxes, np.float32) - np.float32(1)) * np.float32(im_scale) + np.float32(1)
        return self.model.forward(img, rois)
