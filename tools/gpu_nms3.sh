#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_nms_gpu; do
  timeout 1200 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$f.log
done
python tools/nms_diag.py 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e'].get('sync_value',0)),'ms/step',round(d['ms_per_step'],3))
print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()}, 'issued',round(d['roofline']['issued_frac'],3))
PY
python bench.py --config nms_sweep --no-cpu-baseline > gpurun_out/bench_nms_sweep.json 2> gpurun_out/bench_nms.err; echo "nms sweep exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_nms_sweep.json'))
print({k:(round(v['ms'],3), round(v['kept_mean'])) for k,v in d['sweep'].items()})
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_direct_3x3c3_o64 -c 1 -f -o gpurun_out/prof_c11 python tools/nms_diag.py > gpurun_out/ncu_c11.log 2>&1; echo "ncu c11 exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:roi_pool_fused -c 1 -f -o gpurun_out/prof_roi3 python tools/nms_diag.py > gpurun_out/ncu_roi3.log 2>&1; echo "ncu roi exit $?"
