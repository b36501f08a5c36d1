"""a few raw-image steps (uint8 HWC 480x640 -> getImages on the device -> detect + NMS) for the ncu launch list of
get_images_kernel:  ncu --metrics gpu__time_duration.sum -k regex:get_images python tools/prof_raw.py   (GPU box)"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multipathnet_b200 as mpn
from multipathnet_b200 import models, workloads as wl
from multipathnet_b200._lib import CImageTransform
H0, W0, R, NC = 480, 640, 1000, 21
ctx = mpn.Context(0)
spec = models.vgg16_fast_rcnn(NC, seed=1234)
m = mpn.Model(ctx, spec, max_rois=1024, max_h=608, max_w=800)
rng = np.random.default_rng(7)
raw = torch.from_numpy(rng.integers(0, 256, (H0, W0, 3), dtype=np.uint8)).pin_memory()
box = torch.from_numpy(wl.random_boxes(R, H0, W0, 7)).pin_memory()
sc = torch.empty((R, NC)).pin_memory(); bb = torch.empty((R, 4 * NC)).pin_memory()
kp = torch.empty((NC - 1, R), dtype=torch.int32).pin_memory(); ct = torch.empty((NC - 1,), dtype=torch.int32).pin_memory()
tfm = CImageTransform.of(spec.transformer)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t = C.c_int32(-1)
    ctx.check(ctx.lib.mpn_model_detect_nms_submit_u8(m.h, raw.data_ptr(), H0, W0, C.addressof(tfm), 600.0, 1000.0, box.data_ptr(), R, -1.5, 0.3,
                                                     sc.data_ptr(), bb.data_ptr(), kp.data_ptr(), ct.data_ptr(), C.byref(t)), "submit_u8")
    ctx.check(ctx.lib.mpn_model_detect_nms_wait(m.h, t.value), "wait")
print("ok raw", ctx.launch_count)
