"""a few device-resident steps of one bench config, for `ncu -k regex:... python tools/prof_step.py <config> [steps]` (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import multipathnet_b200 as mpn
from multipathnet_b200 import models, workloads as wl
cfg = sys.argv[1] if len(sys.argv) > 1 else "vgg16_frcnn"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
wk = bench.WORKLOADS[cfg]
H, W, R, C = wk["H"], wk["W"], wk["R"], wk["C"]
ctx = mpn.Context(0)
spec = getattr(models, wk["model"])(C, seed=1234, **wk["kw"])
m = mpn.Model(ctx, spec, max_rois=max(R, 1024), max_h=H + 8, max_w=W)
img = torch.from_numpy(wl.transform(wl.raw_image(H, W, 7), spec.transformer)).cuda()
mk = wl.random_boxes if wk["boxes"] == "random" else wl.sharpmask_boxes
boxes = torch.from_numpy(mk(R, H, W, 7)).cuda()
sc = torch.empty((R, C), device="cuda"); bb = torch.empty((R, 4 * C), device="cuda")
kp = torch.empty((C - 1, R), dtype=torch.int32, device="cuda"); ct = torch.empty((C - 1,), dtype=torch.int32, device="cuda")
rec = torch.zeros((steps, mpn.MPN_REC_FLOATS), device="cuda")
m.set_detection_sink(rec, steps, 100)
for i in range(steps):
    m.detect_nms_dev(img, H, W, boxes, R, 1.0, W, H, -1.5, 0.3, sc, bb, kp, ct)
torch.cuda.synchronize()
print("ok", cfg, steps, "launches", ctx.launch_count)
