"""experiment (GPU box): N model replicas on ONE GPU, each with its own mpn_ctx / CUDA stream, images dealt round-robin —
does filling the other replica's layer-boundary and NMS-chain bubbles raise proposals/s?   python tools/two_replicas.py [config] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import multipathnet_b200 as mpn
from multipathnet_b200 import models, workloads as wl
cfg = sys.argv[1] if len(sys.argv) > 1 else "vgg16_frcnn"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
wk = bench.WORKLOADS[cfg]
H, W, R, C = wk["H"], wk["W"], wk["R"], wk["C"]
dev = torch.device("cuda", 0)
for NREP in (1, 2, 3):
    reps = []
    for k in range(NREP):
        st = torch.cuda.Stream(device=dev, priority=0)
        ctx = mpn.Context(0, st.cuda_stream)
        spec = getattr(models, wk["model"])(C, seed=1234, **wk["kw"])
        m = mpn.Model(ctx, spec, max_rois=max(R, 1024), max_h=H + 8, max_w=W)
        img = torch.from_numpy(wl.transform(wl.raw_image(H, W, 7 + k), spec.transformer)).to(dev)
        mk = wl.random_boxes if wk["boxes"] == "random" else wl.sharpmask_boxes
        boxes = torch.from_numpy(mk(R, H, W, 7 + k)).to(dev)
        out = (torch.empty((R, C), device=dev), torch.empty((R, 4 * C), device=dev),
               torch.empty((C - 1, R), dtype=torch.int32, device=dev), torch.empty((C - 1,), dtype=torch.int32, device=dev))
        reps.append((st, ctx, m, img, boxes, out))
    def step(i):
        st, ctx, m, img, boxes, o = reps[i % NREP]
        m.detect_nms_dev(img, H, W, boxes, R, 1.0, W, H, -1.5, 0.3, *o)
    for i in range(6 * NREP):
        step(i)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    print(f"replicas {NREP}: {R * steps / dt:10.0f} proposals/s  {1e3 * dt / steps:.4f} ms/image (wall, {steps} images)", flush=True)
    for st, ctx, m, *_ in reps:
        m.close(); ctx.close()
