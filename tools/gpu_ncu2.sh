#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_nms_gpu test_ops_gpu test_engine_gpu test_model_gpu; do
  timeout 1200 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
python tools/conv_trace.py 2>&1 | cut -c1-330 | tee gpurun_out/conv_trace.log
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e'].get('sync_value',0)),'ms/step',round(d['ms_per_step'],3))
print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()}, 'issued',round(d['roofline']['issued_frac'],3))
PY
tail -n 3 gpurun_out/bench_n1.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:nms_scan_warp_kernel -s 2 -c 1 -f -o gpurun_out/prof_nms3 python tools/nms_diag.py > gpurun_out/ncu_nms3.log 2>&1; echo "ncu nms exit $?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc_kernel -c 1 -f -o gpurun_out/prof_fc6 python tools/nms_diag.py > gpurun_out/ncu_fc6.log 2>&1; echo "ncu fc6 exit $?"
ls -la gpurun_out/*.ncu-rep
