#!/bin/bash
# First GPU call of the next round (one B200, ~6 min of box time):
#   1. the whole GPU suite with the xfail/xpass lines shown: confirms get_images_kernel (tests/test_zz_pending_gpu.py),
#      written after round 1's GPU budget was spent;
#   2. smoke();
#   3. the default bench line, then the same with conv5 on the generic kernel (MPN_TC_R3_MINPIX=2000: the 16 x 8 patches
#      pad the 38 x 50 maps by 29 %, profiles/r01h_layer_efficiency.md), same box, back to back;
#   3b. MultiPathNet with MPN_ROI_NORM_SPLIT=1 against the default ROI kernel;
#   4. getImages on the device: time for a 480 x 640 -> 600 x 800 image (CUDA events around mpn_get_images_dev).
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -rxX -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/summary.txt
grep -E "XPASS|XFAIL|passed|failed" gpurun_out/all_gpu_tests.log | tail -20 | tee -a gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/summary.txt
run() {  # name, env...
  name=$1; shift
  out=$(env "$@" python bench.py --steps 150 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$out" > "gpurun_out/bench_${name}.json"
  python - "$name" "$out" <<'PY' | tee -a gpurun_out/summary.txt
import json,sys
d=json.loads(sys.argv[2])
print(f"{sys.argv[1]:24s} value {d['value']:9.0f}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:9.0f}  clocks {d['clocks']['sm_mhz']}")
PY
}
run default X=1
run conv5_generic MPN_TC_R3_MINPIX=2000
run default_repeat X=1
# MultiPathNet (BASELINE configs[2]): ROI stage with the two-pass normalisation (default off) against the staged kernel
runc() {  # name, env...
  name=$1; shift
  out=$(env "$@" python bench.py --config multipathnet --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$out" > "gpurun_out/bench_${name}.json"
  python - "$name" "$out" <<'PY' | tee -a gpurun_out/summary.txt
import json,sys
d=json.loads(sys.argv[2])
print(f"{sys.argv[1]:24s} value {d['value']:9.0f}  ms/step {d['ms_per_step']:.4f}  roi_pool ms {d['roofline']['by_category_ms_per_step']['roi_pool']:.3f}")
PY
}
runc mpn_default X=1
runc mpn_roi_norm_split MPN_ROI_NORM_SPLIT=1
python - <<'PY' 2>&1 | tee -a gpurun_out/summary.txt
import ctypes as C, numpy as np, torch
import multipathnet_b200 as mpn
from multipathnet_b200 import workloads as wl
from multipathnet_b200._lib import CImageTransform
ctx = mpn.Context(0, None)
im = torch.from_numpy(wl.raw_image(480, 640, 1)).cuda()
h, w, s = C.c_int32(), C.c_int32(), C.c_double()
ctx.lib.mpn_get_images_size(480, 640, 600.0, 1000.0, C.byref(h), C.byref(w), C.byref(s))
out = torch.empty((3, h.value, w.value), device="cuda")
tf = CImageTransform.of("ross")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(3):
    e0.record()
    for _ in range(20):
        ctx.check(ctx.lib.mpn_get_images_dev(ctx.h, im.data_ptr(), 480, 640, C.addressof(tf), h.value, w.value, out.data_ptr()), "get_images")
    e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
gb = (im.numel() + out.numel()) * 4 / 1e9
print(f"get_images 480x640 -> {h.value}x{w.value}: {us:.1f} us, {gb / (us * 1e-6):.0f} GB/s algorithmic")
PY
