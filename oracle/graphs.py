"""CPU restatement of model:forward / ImageDetect:detect / Tester_FRCNN:testOne for a ModelSpec.
TEST INFRASTRUCTURE (see oracle/__init__.py). Dense layers (cudnn.SpatialConvolution, nn.Linear,
nn.SpatialMaxPooling ceil mode, ReLU) are the standard fp32 math via PyTorch-CPU — "parity unpinned"
(third-party nn/cudnn, no reference test holds values); everything else goes through the C
restatement in mpn_oracle.c, which cites the Lua it follows."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ref as O

CONV, MAXPOOL, AVGPOOL, FLATTEN, LRN = 1, 2, 3, 4, 5


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _run_layers(layers, slots, weights):
    for L in layers:
        x = slots[L.in_slot]
        if L.kind == CONV:
            w = _t(weights[L.weight])
            b = _t(weights[L.bias]) if L.bias >= 0 else None
            if x.dim() == 2:                                   # Linear on a flattened tensor
                y = F.linear(x, w.reshape(L.cout, -1), b)
            else:
                g = getattr(L, "groups", 1)
                y = F.conv2d(x, w.reshape(L.cout, L.cin // g, L.kh, L.kw), b, stride=L.stride, padding=L.pad, groups=g)
            if L.residual_slot >= 0:
                y = y + slots[L.residual_slot]
            if L.relu:
                y = F.relu(y)
        elif L.kind == MAXPOOL:
            y = F.max_pool2d(x, (L.kh, L.kw), L.stride, L.pad, ceil_mode=bool(L.ceil_mode))
        elif L.kind == AVGPOOL:
            y = x.mean(dim=(2, 3))                              # avgpool 7 + View (resnet.lua:39)
        elif L.kind == LRN:                                     # CaffeNet norm1/norm2: local_size 5, alpha 1e-4, beta 0.75
            y = F.local_response_norm(x, 5, alpha=1e-4, beta=0.75, k=1.0)
        elif L.kind == FLATTEN:
            y = x.reshape(x.shape[0], -1)                       # (c, ph, pw) order: View(-1):setNumInputDims(3)
        else:
            raise ValueError(L.kind)
        slots[L.out_slot] = y
    return slots


def trunk_forward(spec, image_chw):
    """model:get(1):forward — returns {slot: NCHW tensor}"""
    with torch.no_grad():
        slots = {0: _t(image_chw)[None]}
        return _run_layers(spec.trunk_layers, slots, spec.weights)


def heads_forward(spec, trunk_slots, rois):
    """modules 2..n (eval mode): returns (cls R x C [logits, or probabilities for an integral head],
    bbox R x 4C after BBoxNorm) — what model:forward returns."""
    rois = np.ascontiguousarray(rois, np.float32)
    R = rois.shape[0]
    with torch.no_grad():
        fov = O.foveal(rois).reshape(R, 4, 5) if any(t.region > 0 for t in spec.towers) else None
        feats = []
        for t in spec.towers:
            reg = rois if t.region == 0 else np.ascontiguousarray(fov[:, t.region, :])
            pooled = []
            for slot, scale in t.levels:
                fm = trunk_slots[slot].numpy()
                p = O.roi_pool(fm, reg, t.pooled_w, t.pooled_h, np.float32(scale), spec.roi_variant)
                if t.normalize:                                 # View(-1,nFeat*49) -> Normalize(2) -> View (model_utils.lua:217-220)
                    p = O.l2_normalize(p.reshape(R, -1)).reshape(p.shape)
                pooled.append(p)
            x = np.concatenate(pooled, axis=1)                  # JoinTable(2), conv5|conv4|conv3
            if t.normalize:
                x = x * np.float32(1000.0)                      # MulConstant(1000) (model_utils.lua:240)
            slots = _run_layers(t.layers, {0: _t(x)}, spec.weights)
            feats.append(slots[t.out_slot].reshape(R, -1))
        cat = torch.cat(feats, dim=1)                           # ModelParallelTable concat on dim 2
        cls = []
        for h in spec.cls_heads:
            cls.append(F.linear(cat[:, h.col_begin:h.col_begin + h.col_len], _t(spec.weights[h.weight]), _t(spec.weights[h.bias])))
        hb = spec.bbox_head
        bbox = F.linear(cat[:, hb.col_begin:hb.col_begin + hb.col_len], _t(spec.weights[hb.weight]), _t(spec.weights[hb.bias])).numpy()
        if len(cls) > 1:                                        # integral eval branch: mean of K softmaxes
            c = np.mean(np.stack([O.softmax(c.numpy()) for c in cls], 0), axis=0, dtype=np.float32)
        else:
            c = cls[0].numpy()
        if spec.has_bbox_norm:
            bbox = O.bbox_norm(bbox, spec.bbox_mean, spec.bbox_std)
        return c, bbox


def detect(spec, image_chw, boxes, im_scale):
    """ImageDetect:detect after getImages (ImageDetect.lua:161-192): (scores R x C, bboxes R x 4C)."""
    rois = O.project_rois(boxes, np.float32(im_scale))
    ts = trunk_forward(spec, image_chw)
    cls, bbox = heads_forward(spec, ts, rois)
    bboxes = O.convert_from(bbox, boxes)
    scores = cls if (spec.no_softmax or len(spec.cls_heads) > 1) else O.softmax(cls)
    return scores, bboxes


def test_one(spec, image_chw, boxes, im_scale, W0, H0, score_thresh=-1.5, nms_thr=0.3, nms_fn=None):
    """Tester_FRCNN:testOne (Tester_FRCNN.lua:72-117): detect, clamp, per-class gather + NMS.
    Returns (scores, clamped bboxes, [keep row indices per foreground class])."""
    nms_fn = nms_fn or O.nms
    scores, bboxes = detect(spec, image_chw, boxes, im_scale)
    bboxes = O.clamp_boxes(bboxes, W0, H0)
    keeps = []
    for j in range(1, scores.shape[1]):
        sel = np.nonzero(scores[:, j] > score_thresh)[0]
        sb = np.concatenate([bboxes[sel, 4 * j:4 * j + 4], scores[sel, j:j + 1]], 1).astype(np.float32)
        k = nms_fn(sb, nms_thr)
        keeps.append(sel[k].astype(np.int32))
    return scores, bboxes, keeps
