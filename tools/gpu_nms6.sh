#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_nms_gpu.py tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
for c in resnet50; do
  python bench.py --config $c --steps 30 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?"
done
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --config nms_sweep --no-cpu-baseline > gpurun_out/bench_nms_sweep.json 2> gpurun_out/bench_nms.err; echo "nms sweep exit $?"
python - <<'PY'
import json
for c in ('n1','resnet50'):
    d=json.load(open(f'gpurun_out/bench_{c}.json'))
    print(c,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3))
    print('  ',{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()}, 'issued',round(d['roofline']['issued_frac'],3))
d=json.load(open('gpurun_out/bench_nms_sweep.json')); print({k:(round(v['ms'],3), round(v['kept_mean'])) for k,v in d['sweep'].items()})
PY
