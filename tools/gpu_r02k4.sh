#!/bin/bash
# round 2, call K4 (one B200): roi_pool_ring_kernel (roi_impl 5: persistent warp-specialised bulk-copy ring) — parity tests with a
# short timeout first (a hang must not take the box), then same-box A/B against the cluster kernel on the three configs, ncu of both on cfg 3
mkdir -p gpurun_out; S=gpurun_out/summary_k4.txt; : > $S
timeout 300 python -m pytest tests/test_roi_product_gpu.py -q -m gpu -x -p no:cacheprovider -k "small" > gpurun_out/k4_tests_small.log 2>&1
rc=$?; echo "small roi tests exit $rc" | tee -a $S; tail -3 gpurun_out/k4_tests_small.log | tee -a $S
if [ $rc -ne 0 ]; then echo "STOP: small tests failed" | tee -a $S; exit 0; fi
timeout 900 python -m pytest tests/test_roi_product_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/k4_tests_roi.log 2>&1
echo "roi tests exit $?" | tee -a $S; tail -3 gpurun_out/k4_tests_roi.log | tee -a $S
run() { name=$1; shift; timeout 600 env "$@" > gpurun_out/k4_bench_$name.json 2> gpurun_out/k4_bench_$name.err; echo "bench $name exit $?" >> $S; }
run n1 X=1 python bench.py --no-cpu-baseline
run n1_ring MPN_ROI_IMPL=5 python bench.py --no-cpu-baseline
run mpn X=1 python bench.py --no-cpu-baseline --config multipathnet --steps 60
run mpn_ring MPN_ROI_IMPL=5 python bench.py --no-cpu-baseline --config multipathnet --steps 60
run resnet50 X=1 python bench.py --no-cpu-baseline --config resnet50 --steps 40
run resnet50_ring MPN_ROI_IMPL=5 python bench.py --no-cpu-baseline --config resnet50 --steps 40
python - <<'PY' 2>&1 | tee -a $S
import json
for c in ('n1', 'n1_ring', 'mpn', 'mpn_ring', 'resnet50', 'resnet50_ring'):
    try:
        d = json.load(open(f'gpurun_out/k4_bench_{c}.json'))
        r = d['roofline']; b = r['by_category_ms_per_step']
        print(f"{c:16s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} p50 {d['ms_per_image_p50']:.4f} e2e {d['e2e']['value']:9.0f} tc {b['conv_gemm_tc']:.4f} roi {b['roi_pool']:.4f} nms {b['nms']:.4f} roi frac {r['roi_pool']['frac']:.3f} clk {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(c, 'ERR', e)
PY
for k in "vgg16_frcnn:roi_pool_ring:bulk_cfg2" "multipathnet:roi_pool_ring:bulk_cfg3"; do
  IFS=: read cfg kn tag <<< "$k"
  MPN_ROI_IMPL=5 timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$kn" -s 1 -c 1 -f -o gpurun_out/r02k4_$tag python tools/prof_step.py $cfg 3 > gpurun_out/ncu_r02k4_$tag.log 2>&1; echo "ncu $tag exit $?" | tee -a $S
done
