#!/bin/bash
# GPU box: full GPU suite + all bench configs + ncu evidence
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_nms_gpu test_ops_gpu test_engine_gpu test_model_gpu; do
  timeout 1200 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
for c in multipathnet resnet50 nms_sweep; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"
done
python - <<'PY'
import json
for c in ('n1','multipathnet','resnet50'):
    try:
        d=json.load(open(f'gpurun_out/bench_{c}.json'))
        print(c,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'p50',round(d['ms_per_image_p50'],3),'launches',d['gpu_launches'])
        print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()})
        print('   tc achieved',round(d['roofline']['achieved'],1),'frac',round(d['roofline']['frac'],3),'issued',round(d['roofline']['issued_frac'],3),'roi GB/s',round(d['roofline']['roi_pool']['achieved'] or 0),'roi frac',d['roofline']['roi_pool']['frac'], d['clocks'], d.get('cpu_baseline'))
    except Exception as e: print(c,'ERR',e)
try:
    d=json.load(open('gpurun_out/bench_nms_sweep.json'))
    for n,v in d['sweep'].items(): print('nms',n,{k:(round(x,3) if isinstance(x,float) else x) for k,x in v.items()})
except Exception as e: print('nms ERR',e)
PY
tail -3 gpurun_out/bench_*.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 40 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
for x in list(csv.DictReader(lines))[:36]:
    n=re.sub(r'\(.*','',x['Kernel Name']).replace('<unnamed>::','').replace('void ','')
    print(x['ID'], n[:34], x['Grid Size'], x['Metric Value'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_scan_small -s 4 -c 1 -o gpurun_out/prof_nms_small \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_nms.log 2>&1; echo "ncu nms exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 68 -c 4 -o gpurun_out/prof_tc2 \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fused -s 4 -c 1 -o gpurun_out/prof_roi2 \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_roi.log 2>&1; echo "ncu roi exit $?"
