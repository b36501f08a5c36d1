#!/bin/bash
# round 2, call F (one B200): the whole GPU suite (no -x), both bench arms, ROI kernel A/B (st.async exchange, 5 CTAs/SM), ncu of the
# ROI stage + the w16 fc6 kernel + get_images_kernel, launch lists of the three configs, compute-sanitizer on the smoke path
mkdir -p gpurun_out; : > gpurun_out/summary_f.txt
S=gpurun_out/summary_f.txt
timeout 2400 python -m pytest tests -q -m gpu -rxXs -p no:cacheprovider --durations=8 > gpurun_out/f_all_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a $S; tail -12 gpurun_out/f_all_gpu_tests.log | tee -a $S
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $S
python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench n1 exit $?" | tee -a $S
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "bench reference exit $?" | tee -a $S
run() { name=$1; cfg=$2; steps=$3; shift 3
  env "$@" python bench.py --config $cfg --steps $steps --no-cpu-baseline > gpurun_out/f_bench_$name.json 2> gpurun_out/f_bench_$name.err; echo "bench $name exit $?" >> $S; }
run n1_s20 vgg16_frcnn 20 X=1
run n1_minb5 vgg16_frcnn 200 MPN_ROI_MINB=5
run n1_roi3 vgg16_frcnn 200 MPN_ROI_IMPL=3
run n1_roi1 vgg16_frcnn 200 MPN_ROI_IMPL=1
run mpn multipathnet 60 X=1
run mpn_minb5 multipathnet 60 MPN_ROI_MINB=5
run mpn_roi3 multipathnet 60 MPN_ROI_IMPL=3
run mpn_roi2 multipathnet 60 MPN_ROI_IMPL=2
run resnet50 resnet50 40 X=1
run resnet50_minb5 resnet50 40 MPN_ROI_MINB=5
python bench.py --config nms_sweep > gpurun_out/f_bench_nms_sweep.json 2> gpurun_out/f_bench_nms_sweep.err; echo "bench nms_sweep exit $?" | tee -a $S
python - <<'PY' 2>&1 | tee -a $S
import json
for c in ('n1', 'n1_s20', 'n1_minb5', 'n1_roi3', 'n1_roi1', 'mpn', 'mpn_minb5', 'mpn_roi3', 'mpn_roi2', 'resnet50', 'resnet50_minb5'):
    try:
        d = json.load(open(f'gpurun_out/f_bench_{c}.json'))
        r = d['roofline']; b = r['by_category_ms_per_step']
        print(f"{c:15s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} p50 {d['ms_per_image_p50']:.4f} e2e {d['e2e']['value']:9.0f} "
              f"tc {b['conv_gemm_tc']:.4f} roi {b['roi_pool']:.4f} nms {b['nms']:.4f} roi frac {r['roi_pool']['frac']:.3f} tc frac {r['frac']:.3f} clk {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(c, 'ERR', e)
try:
    d = json.load(open('gpurun_out/f_bench_reference.json')); print('reference', d['value'], d['cpu_baseline'])
except Exception as e: print('reference ERR', e)
PY
for k in "vgg16_frcnn:roi_pool_cluster:roi_cfg2" "multipathnet:roi_pool_cluster:roi_cfg3" "vgg16_frcnn:conv_gemm_tc_kernel<240:fc6_w16"; do
  IFS=: read cfg kn tag <<< "$k"
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$kn" -s 1 -c 1 -f -o gpurun_out/r02f_$tag python tools/prof_step.py $cfg 3 > gpurun_out/ncu_r02f_$tag.log 2>&1; echo "ncu $tag exit $?" | tee -a $S
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_cfg2.csv python tools/prof_step.py vgg16_frcnn 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_mpn.csv python tools/prof_step.py multipathnet 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02f_launches_resnet50.csv python tools/prof_step.py resnet50 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:get_images --csv --log-file gpurun_out/r02f_launches_getimages.csv python tools/prof_raw.py 4 > gpurun_out/prof_raw.log 2>&1
echo "launch lists done" | tee -a $S
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1
echo "memcheck smoke exit $?" | tee -a $S; tail -3 gpurun_out/r02_sanitizer_memcheck_smoke.log | tee -a $S
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1
echo "racecheck smoke exit $?" | tee -a $S; tail -3 gpurun_out/r02_sanitizer_racecheck_smoke.log | tee -a $S
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_post_gpu.py tests/test_ops_gpu.py tests/test_roi_product_gpu.py -q -m gpu -x -p no:cacheprovider -k "not full_size" > gpurun_out/r02_sanitizer_memcheck_ops.log 2>&1
echo "memcheck ops exit $?" | tee -a $S; tail -3 gpurun_out/r02_sanitizer_memcheck_ops.log | tee -a $S
