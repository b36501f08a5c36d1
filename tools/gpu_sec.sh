#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
for c in multipathnet resnet50; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_$c.csv \
   python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$c.log 2>&1; echo "ncu launches $c exit $?"
python - $c <<'PY'
import csv,re,sys,collections
c=sys.argv[1]
lines=[l for l in open(f'gpurun_out/launches_{c}.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg=collections.OrderedDict()
for x in rows:
    n=re.sub(r'\(.*','',x['Kernel Name']).replace('<unnamed>::','').replace('void ','')
    k=(n[:40], x['Grid Size'])
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(x['Metric Value'].replace(',',''))
tot=sum(v[1] for v in agg.values())
print(c,'total us in window',round(tot/1e3,1))
for k,v in sorted(agg.items(), key=lambda t:-t[1][1])[:22]:
    print(f"  {k[0]:40s} {k[1]:16s} x{v[0]:3d}  {v[1]/1e3:9.1f} us")
PY
done
