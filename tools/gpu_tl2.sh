#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_engine_gpu test_model_gpu; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
python tools/timeline.py 2>&1 | tee gpurun_out/timeline.log | cut -c1-60,150-230
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
