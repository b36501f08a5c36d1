#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_engine_gpu test_model_gpu; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
run() { name=$1; shift
  out=$(env "$@" python bench.py --steps 150 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$name" "$out" <<'PY'
import json,sys
d=json.loads(sys.argv[2])
print(f"{sys.argv[1]:24s} value {d['value']:9.0f}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:9.0f}  clocks {d['clocks']['sm_mhz']}")
PY
}
run "b prefetch on" X=1
run "MPN_TC_BPREFETCH=0" MPN_TC_BPREFETCH=0
run "b prefetch on (rep)" X=1
run "MPN_TC_BPREFETCH=0 (rep)" MPN_TC_BPREFETCH=0
