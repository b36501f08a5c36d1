#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nms_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3))
print({k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()})
PY
timeout 300 python bench.py --config nms_sweep 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
for n,v in d['sweep'].items(): print('nms',n,round(v['ms'],3),'ms', round(v['pair_ious_per_s']/1e9),'G pair/s kept',round(v['kept_mean']))
"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms_ -s 12 -c 6 --csv --log-file gpurun_out/launches_nms.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_nms.csv | cut -d, -f5,15 | tail -7
