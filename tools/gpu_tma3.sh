#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_engine_gpu test_model_gpu test_ops_gpu; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
run() { name=$1; cfg=$2; shift 2
  out=$(env "$@" python bench.py --config $cfg --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$name" "$out" <<'PY'
import json,sys
d=json.loads(sys.argv[2])
print(f"{sys.argv[1]:34s} value {d['value']:9.0f}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:9.0f}  clocks {d['clocks']['sm_mhz']}")
PY
}
run "resnet50 tma on" resnet50 X=1
run "resnet50 MPN_TC_TMA_STORE=0" resnet50 MPN_TC_TMA_STORE=0
run "multipathnet tma on" multipathnet X=1
run "multipathnet MPN_TC_TMA_STORE=0" multipathnet MPN_TC_TMA_STORE=0
run "vgg16_frcnn tma on" vgg16_frcnn X=1
