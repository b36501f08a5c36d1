#!/bin/bash
# round 2, call I (one B200): ROI kernel third pass (coincident block positions loaded once; evict-first stores for pooled tensors >> L2):
# ROI tests, A/B benches, ncu of the ROI stage of cfg 2 / cfg 3
mkdir -p gpurun_out; S=gpurun_out/summary_i.txt; : > $S
timeout 1200 python -m pytest tests/test_roi_product_gpu.py tests/test_model_gpu.py tests/test_ops_gpu.py -q -m gpu -rs -p no:cacheprovider > gpurun_out/i_tests.log 2>&1
echo "tests exit $?" | tee -a $S; tail -4 gpurun_out/i_tests.log | tee -a $S
run() { name=$1; shift; env "$@" > gpurun_out/i_bench_$name.json 2> gpurun_out/i_bench_$name.err; echo "bench $name exit $?" >> $S; }
run n1 X=1 python bench.py --no-cpu-baseline
run mpn X=1 python bench.py --no-cpu-baseline --config multipathnet --steps 60
run mpn_nostcs MPN_ROI_STCS=0 python bench.py --no-cpu-baseline --config multipathnet --steps 60
run mpn_minb5 MPN_ROI_MINB=5 python bench.py --no-cpu-baseline --config multipathnet --steps 60
run resnet50 X=1 python bench.py --no-cpu-baseline --config resnet50 --steps 40
run resnet50_nostcs MPN_ROI_STCS=0 python bench.py --no-cpu-baseline --config resnet50 --steps 40
python - <<'PY' 2>&1 | tee -a $S
import json
for c in ('n1', 'mpn', 'mpn_nostcs', 'mpn_minb5', 'resnet50', 'resnet50_nostcs'):
    try:
        d = json.load(open(f'gpurun_out/i_bench_{c}.json'))
        r = d['roofline']; b = r['by_category_ms_per_step']
        print(f"{c:16s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} p50 {d['ms_per_image_p50']:.4f} e2e {d['e2e']['value']:9.0f} tc {b['conv_gemm_tc']:.4f} roi {b['roi_pool']:.4f} nms {b['nms']:.4f} roi frac {r['roi_pool']['frac']:.3f} clk {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(c, 'ERR', e)
PY
for k in "vgg16_frcnn:roi_pool_cluster:roi_cfg2" "multipathnet:roi_pool_cluster:roi_cfg3"; do
  IFS=: read cfg kn tag <<< "$k"
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$kn" -s 1 -c 1 -f -o gpurun_out/r02i_$tag python tools/prof_step.py $cfg 3 > gpurun_out/ncu_r02i_$tag.log 2>&1; echo "ncu $tag exit $?" | tee -a $S
done
