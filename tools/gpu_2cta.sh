#!/bin/bash
# GPU box: bring-up of the cta_group::2 engine (both engines), then tests + bench + launch list
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
echo "== first light, default engine (cta_group::2 where >= 2 m-tiles)"; timeout 600 python tools/first_light.py 2>&1 | cut -c1-230
cp gpurun_out/first_light.log gpurun_out/first_light_cg2.log
if grep -q -E "TIMEOUT|NO RESULT|bad_frac" gpurun_out/first_light_cg2.log; then
  echo "!! cta_group::2 engine failed bring-up: continuing with MPN_TC_CTA_GROUP=1"; export MPN_TC_CTA_GROUP=1
fi
for f in test_nms_gpu test_engine_gpu test_model_gpu; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
MPN_TC_CTA_GROUP=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_cg1.json 2> gpurun_out/bench_n1_cg1.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_n1.json','gpurun_out/bench_n1_cg1.json'):
    try:
        d=json.load(open(f))
        print(f,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'launches',d['gpu_launches'])
        print({k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()})
        print('tc achieved',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3),'issued',round(d['roofline']['issued_frac'],3), d['clocks'])
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 40 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
for x in list(csv.DictReader(lines))[:34]:
    n=re.sub(r'\(.*','',x['Kernel Name']).replace('<unnamed>::','').replace('void ','')
    print(x['ID'], n[:34], x['Grid Size'], x['Metric Value'])
PY
