"""tie structure of the default bench workload's scores + time of the NMS stage on exactly those boxes (GPU box)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipathnet_b200 as mpn
from multipathnet_b200 import models, workloads as wl
ctx = mpn.Context(0)
C, H, W, R = 21, 600, 800, 1000
spec = models.vgg16_fast_rcnn(C, seed=1234)
m = mpn.Model(ctx, spec, max_rois=R, max_h=H, max_w=W)
img = wl.transform(wl.raw_image(H, W, 7), "ross"); boxes = wl.random_boxes(R, H, W, 100)
scores, bboxes, keep = m.detect_nms(img, boxes, 1.0, W, H, -1.5, 0.3)
print("score range", scores.min(), scores.max(), "kept/class", [len(k) for k in keep])
for c in range(1, C):
    s = scores[:, c]
    u, cnt = np.unique(s, return_counts=True)
    g = cnt[cnt > 1]
    print(f"class {c}: distinct {len(u)} tie groups {len(g)} members {int(g.sum())} largest {int(g.max()) if len(g) else 0} zeros {(s == 0).sum()} ones {(s == 1).sum()}")
# time the NMS stage alone on these scored boxes
sb = np.zeros((C - 1, R, 5), np.float32)
for c in range(1, C):
    sb[c - 1, :, :4] = bboxes[:, 4 * c:4 * c + 4]; sb[c - 1, :, 4] = scores[:, c]
offs = [i * R for i in range(C)]
for rep in range(3):
    ctx.profile_begin()
    for _ in range(20):
        ctx.nms_batched(sb, offs, 0.3)
    p = ctx.profile_end()
    print({k: round(v[0] / 20, 4) for k, v in p.items() if v[1]})
