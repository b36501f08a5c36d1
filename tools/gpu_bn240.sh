#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
python - <<'PY'
import sys; sys.path.insert(0,'.')
import multipathnet_b200 as mpn
ctx = mpn.Context(0)
for (M,N,K) in [(1000,4096,25088),(1000,4096,4096),(2000,4096,25088),(300,4096,25088)]:
    ms,bn,cg,sk = ctx.gemm_bench(M,N,K,20)
    print(M,N,K,'ms',round(ms,4),'BN',bn,'CG',cg,'splitk',sk,'issued_TF',round(2.0*M*N*K*3/(ms*1e-3)/1e12))
PY
for f in test_engine_gpu test_model_gpu; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['e2e'].get('sync_value',0)),'ms/step',round(d['ms_per_step'],3))
print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()}, 'issued',round(d['roofline']['issued_frac'],3))
PY
