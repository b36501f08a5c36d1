#!/bin/bash
# round-end evidence run: tests, smoke, benches of every config, launch list, ncu --set full of the top kernels
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_nms_gpu test_ops_gpu test_engine_gpu test_model_gpu; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$f.log
done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
for c in multipathnet resnet50; do
  python bench.py --config $c --steps 30 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"
done
python bench.py --config nms_sweep --no-cpu-baseline > gpurun_out/bench_nms_sweep.json 2> gpurun_out/bench_nms.err; echo "nms sweep exit $?"
python - <<'PY'
import json
for c in ('n1','multipathnet','resnet50'):
    try:
        d=json.load(open(f'gpurun_out/bench_{c}.json'))
        print(c,'value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e'].get('sync_value',0)),'ms/step',round(d['ms_per_step'],3), d['clocks'])
        print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()}, 'issued',round(d['roofline']['issued_frac'],3), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(c,'ERR',e)
d=json.load(open('gpurun_out/bench_nms_sweep.json')); print({k:(round(v['ms'],3), round(v['kept_mean'])) for k,v in d['sweep'].items()})
print(open('gpurun_out/bench_ref.json').read()[:300])
PY
python tools/conv_trace.py 2>&1 | cut -c1-330 > gpurun_out/conv_trace.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 40 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/launches.csv') if not l.startswith('==')]
for x in list(csv.DictReader(lines))[:30]:
    n=re.sub(r'\(.*','',x['Kernel Name']).replace('<unnamed>::','').replace('void ','')
    print(x['ID'], n[:34], x['Grid Size'], x['Metric Value'])
PY
for k in "conv_gemm_tc_kernel:fc6:0" "conv3x3_tc_kernel:r3:7" "roi_pool_fused:roi:0" "nms_scan_warp_kernel:nmswalk:0" "conv1_tc_kernel:conv1:0" "nms_mask_kernel:nmsmask:0"; do
  IFS=: read kn tag skip <<< "$k"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$kn -s $skip -c 1 -f -o gpurun_out/final_$tag python tools/nms_diag.py > gpurun_out/ncu_final_$tag.log 2>&1; echo "ncu $tag exit $?"
done
