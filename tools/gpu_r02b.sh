#!/bin/bash
# round 2, call B (one B200): w16 engine tests, product-ROI parity, A/B benches (roi_impl, fc_w16), ncu of the ROI stage + fc6
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_roi_product_gpu.py tests/test_post_gpu.py -q -m gpu -x -rs -p no:cacheprovider > gpurun_out/tests_b1.log 2>&1
echo "engine+roi+post tests exit $?" | tee -a gpurun_out/summary.txt; tail -4 gpurun_out/tests_b1.log | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x -rs -s -p no:cacheprovider > gpurun_out/tests_b2.log 2>&1
echo "model tests exit $?" | tee -a gpurun_out/summary.txt; grep -E "rel err|passed|failed|Error" gpurun_out/tests_b2.log | tail -6 | tee -a gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/summary.txt
run() { name=$1; cfg=$2; steps=$3; shift 3
  env "$@" python bench.py --config $cfg --steps $steps --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name exit $?" >> gpurun_out/summary.txt; }
run n1 vgg16_frcnn 100 X=1
run n1_w16off vgg16_frcnn 100 MPN_FC_W16=0
run n1_roi1 vgg16_frcnn 100 MPN_ROI_IMPL=1
run n1_again vgg16_frcnn 100 X=1
run mpn multipathnet 40 X=1
run mpn_w16off multipathnet 40 MPN_FC_W16=0
run mpn_roi2 multipathnet 40 MPN_ROI_IMPL=2
run resnet50 resnet50 30 X=1
run resnet50_roi1 resnet50 30 MPN_ROI_IMPL=1
python - <<'PY' 2>&1 | tee -a gpurun_out/summary.txt
import json
for c in ('n1', 'n1_w16off', 'n1_roi1', 'n1_again', 'mpn', 'mpn_w16off', 'mpn_roi2', 'resnet50', 'resnet50_roi1'):
    try:
        d = json.load(open(f'gpurun_out/bench_{c}.json'))
        r = d['roofline']; b = r['by_category_ms_per_step']
        print(f"{c:14s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} p50 {d['ms_per_image_p50']:.4f} e2e {d['e2e']['value']:9.0f} "
              f"tc ms {b['conv_gemm_tc']:.4f} roi ms {b['roi_pool']:.4f} elt {b['elementwise']:.4f} nms {b['nms']:.4f} roi frac {r['roi_pool']['frac']:.3f} tc frac {r['frac']:.3f} clk {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(c, 'ERR', e)
PY
for k in "vgg16_frcnn:roi_pool_cluster:roi_cfg2:1" "multipathnet:roi_pool_cluster:roi_cfg3:1" "vgg16_frcnn:conv_gemm_tc_kernel:fc6w16:4"; do
  IFS=: read cfg kn tag skip <<< "$k"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kn -s $skip -c 1 -f -o gpurun_out/r02b_$tag python tools/prof_step.py $cfg 3 > gpurun_out/ncu_r02b_$tag.log 2>&1; echo "ncu $tag exit $?" | tee -a gpurun_out/summary.txt
done
