"""fbcoco.ImageDetect mirror (ImageDetect.lua) over the C ABI."""
from __future__ import annotations

import math

import numpy as np

from ._lib import Model
from .modules import ImageTransformer


def _scale_axis(src: np.ndarray, dst_len: int) -> np.ndarray:
    """One pass of image.scale's 'bilinear' mode along the LAST axis, fp32 step by step (torch `image` package,
    generic/image.c scaleLinear_rowcol as recalled; third-party, unpinned => "parity unpinned"):
    a longer axis is corner-aligned linear interpolation, a shorter one an area average, an equal one a copy."""
    f32 = np.float32
    src_len = src.shape[-1]
    if dst_len == src_len:
        return src.copy()
    out = np.empty(src.shape[:-1] + (dst_len,), f32)
    if dst_len > src_len:
        if src_len == 1:
            out[...] = src[..., :1]
            return out
        scale = f32(src_len - 1) / f32(dst_len - 1)
        sf = np.arange(dst_len - 1, dtype=f32) * scale
        si = sf.astype(np.int64)
        sf = sf - si.astype(f32)
        out[..., :-1] = (f32(1) - sf) * src[..., si] + sf * src[..., si + 1]
        out[..., -1] = src[..., -1]
        return out
    scale = f32(src_len) / f32(dst_len)
    s0_i, s0_f = 0, f32(0)
    for di in range(dst_len):                       # the window state is carried from sample to sample, as in the C code
        s1_f = f32(di + 1) * scale
        s1_i = int(s1_f)
        s1_f = s1_f - f32(s1_i)
        acc = (f32(1) - s0_f) * src[..., s0_i]
        n = f32(1) - s0_f
        for si in range(s0_i + 1, s1_i):
            acc = acc + src[..., si]
            n = n + f32(1)
        if s1_i < src_len:
            acc = acc + s1_f * src[..., s1_i]
            n = n + s1_f
        out[..., di] = acc / n
        s0_i, s0_f = s1_i, s1_f
    return out


def _image_scale(im: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """image.scale(im, w, h): rows to the new width first, then the columns of that temporary (scaleBilinear)."""
    im = np.ascontiguousarray(im, np.float32)
    tmp = _scale_axis(im, out_w)                                        # C x H x w
    return np.ascontiguousarray(_scale_axis(tmp.transpose(0, 2, 1), out_h).transpose(0, 2, 1))


def _get_images_size(h: int, w: int, scale: float, max_size: float):
    """ImageDetect.lua:31-39: im_scale (a Lua double) and the size image.scale allocates (numbers truncated to long)."""
    smin, smax = min(h, w), max(h, w)
    im_scale = scale / smin
    if math.floor(im_scale * smax + 0.5) > max_size:                     # torch.round, not Python's banker's rounding
        im_scale = max_size / smax
    return int(h * im_scale), int(w * im_scale), float(im_scale)


class ImageDetect:
    def __init__(self, model: Model, transformer: ImageTransformer, scale=None, max_size=None, on_device: bool = False):
        if model is None:
            raise ValueError("must provide model!")           # ImageDetect.lua:13
        if transformer is None:
            raise ValueError("must provide transformer!")     # ImageDetect.lua:14
        self.model = model
        self.image_transformer = transformer
        self.scale = list(scale) if scale else [600]
        self.max_size = max_size or 1000
        # on_device: getImages runs in the library (mpn_model_trunk_image: transformer + image.scale in one kernel on the
        # raw image) instead of on the host in numpy; same arithmetic, see csrc/image_scale.cuh
        self.on_device = bool(on_device)
        if len(self.scale) != 1:
            # project_im_rois' multi-scale branch never fills rois (ImageDetect.lua:57-65): single scale only
            raise ValueError("only a single test scale is functional in the reference")

    def getImages(self, im):
        """ImageDetect.lua:22-52: transformer, scale to self.scale (capped by max_size). Host side (numpy)."""
        im = self.image_transformer.forward(im)
        out_h, out_w, im_scale = _get_images_size(im.shape[1], im.shape[2], self.scale[0], self.max_size)
        return _image_scale(im, out_h, out_w), im_scale

    def detect(self, im, boxes, min_images=None, recompute_features=True):
        """-> (scores R x C float32, bboxes R x 4C float32) in original-image coordinates.
        `min_images` (DataParallelTable width) only replicated the image in the reference; ignored."""
        boxes = np.ascontiguousarray(boxes, np.float32)
        if boxes.ndim != 2 or boxes.shape[1] != 4:
            raise ValueError("boxes must be R x 4 [x1,y1,x2,y2]")
        if recompute_features and self.on_device:
            self._im_scale, _h, _w = self.model.trunk_image(im, self.image_transformer.kind, self.scale[0], self.max_size)
            return self.model.detect(None, boxes, self._im_scale, False)
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
