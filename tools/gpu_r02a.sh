#!/bin/bash
# round 2, call A (one B200): the new parity tests, benches with the three ROI implementations, ncu of the ROI stage
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1700 python -m pytest tests -q -m gpu -x -rxXs -p no:cacheprovider --durations=8 > gpurun_out/all_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/all_gpu_tests.log | tee -a gpurun_out/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/summary.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_s20.json 2> gpurun_out/bench_n1_s20.err; echo "bench s20 exit $?" | tee -a gpurun_out/summary.txt
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
for impl in 0 1 2; do
  MPN_ROI_IMPL=$impl python bench.py --config multipathnet --steps 40 --no-cpu-baseline > gpurun_out/bench_mpn_roi$impl.json 2> gpurun_out/bench_mpn_roi$impl.err
  echo "bench mpn roi_impl $impl exit $?" | tee -a gpurun_out/summary.txt
done
MPN_ROI_IMPL=1 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/bench_n1_roi1.json 2> gpurun_out/bench_n1_roi1.err
for impl in 0 1; do
  MPN_ROI_IMPL=$impl python bench.py --config resnet50 --steps 30 --no-cpu-baseline > gpurun_out/bench_resnet50_roi$impl.json 2> gpurun_out/bench_resnet50_roi$impl.err
done
python - <<'PY' 2>&1 | tee -a gpurun_out/summary.txt
import json
for c in ('n1_s20', 'n1', 'n1_roi1', 'mpn_roi0', 'mpn_roi1', 'mpn_roi2', 'resnet50_roi0', 'resnet50_roi1'):
    try:
        d = json.load(open(f'gpurun_out/bench_{c}.json'))
        r = d['roofline']
        print(f"{c:14s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} p50 {d['ms_per_image_p50']:.4f} e2e {d['e2e']['value']:9.0f} coll_ms {d['collective']['ms']:.4f} "
              f"roi ms {r['by_category_ms_per_step']['roi_pool']:.4f} roi frac {r['roi_pool']['frac']:.3f} tc frac {r['frac']:.3f} clk {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(c, 'ERR', e)
PY
for k in "vgg16_frcnn:roi_pool_cluster:roi_cfg2" "multipathnet:roi_pool_cluster:roi_cfg3" "multipathnet:maxpyr:pyr_cfg3" "vgg16_frcnn:pack_detections:pack"; do
  IFS=: read cfg kn tag <<< "$k"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kn -s 1 -c 1 -f -o gpurun_out/r02a_$tag python tools/prof_step.py $cfg 3 > gpurun_out/ncu_r02a_$tag.log 2>&1; echo "ncu $tag exit $?" | tee -a gpurun_out/summary.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_launches_mpn.csv python tools/prof_step.py multipathnet 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_launches_cfg2.csv python tools/prof_step.py vgg16_frcnn 2 > /dev/null 2>&1
