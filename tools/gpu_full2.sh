#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_nms_gpu test_ops_gpu test_engine_gpu test_model_gpu; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
for c in multipathnet resnet50; do
  python bench.py --config $c --steps 30 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"
done
python - <<'PY'
import json
for c in ('n1','multipathnet','resnet50'):
    try:
        d=json.load(open(f'gpurun_out/bench_{c}.json'))
        print(c,'value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e'].get('sync_value',0)),'ms/step',round(d['ms_per_step'],3))
        print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()}, 'issued',round(d['roofline']['issued_frac'],3))
    except Exception as e: print(c,'ERR',e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 40 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
for x in list(csv.DictReader(lines))[:40]:
    n=re.sub(r'\(.*','',x['Kernel Name']).replace('<unnamed>::','').replace('void ','')
    print(x['ID'], n[:34], x['Grid Size'], x['Metric Value'])
PY
