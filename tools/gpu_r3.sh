#!/bin/bash
# GPU box: bring-up of the 3x3 A-reuse kernel, then tests + bench + launch list
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
echo "== first light (default engines: R3 for 3x3 convs)"; timeout 600 python tools/first_light.py 2>&1 | cut -c1-200 | grep conv
cp gpurun_out/first_light.log gpurun_out/first_light_r3.log
if grep conv gpurun_out/first_light_r3.log | grep -q -E "TIMEOUT|NO RESULT|bad_frac"; then
  echo "!! R3 kernel failed bring-up: continuing with MPN_TC_R3=0"; export MPN_TC_R3=0
fi
for f in test_engine_gpu test_model_gpu; do
  timeout 1200 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -6 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
MPN_TC_R3=0 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_nor3.json 2> gpurun_out/bench_n1_nor3.err
python bench.py --config multipathnet --steps 20 --warmup 3 > gpurun_out/bench_multipathnet.json 2> gpurun_out/bench_multipathnet.err
python - <<'PY'
import json
for c in ('n1','n1_nor3','multipathnet'):
    try:
        d=json.load(open(f'gpurun_out/bench_{c}.json'))
        print(c,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3))
        print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()}, 'issued',round(d['roofline']['issued_frac'],3))
    except Exception as e: print(c,'ERR',e)
PY
for f in gpurun_out/bench_n1.err; do tail -n 3 $f; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 30 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
for x in list(csv.DictReader(lines))[:30]:
    n=re.sub(r'\(.*','',x['Kernel Name']).replace('<unnamed>::','').replace('void ','')
    print(x['ID'], n[:34], x['Grid Size'], x['Metric Value'])
PY
