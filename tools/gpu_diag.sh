#!/bin/bash
mkdir -p gpurun_out
python tools/nms_diag.py 2>&1 | tee gpurun_out/nms_diag.log
python tools/conv_trace.py 2>&1 | cut -c1-330 | tee gpurun_out/conv_trace.log
for f in test_engine_gpu test_ops_gpu; do
  timeout 1200 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?"; tail -4 gpurun_out/$f.log
done
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3))
print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()}, 'issued',round(d['roofline']['issued_frac'],3))
PY
