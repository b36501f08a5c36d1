#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_model_gpu test_engine_gpu; do
  timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider -x > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
run() { name=$1; shift
  out=$(env "$@" python bench.py --steps 150 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$name" "$out" <<'PY'
import json,sys
d=json.loads(sys.argv[2])
print(f"{sys.argv[1]:24s} value {d['value']:9.0f}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:9.0f}  launches/step {d['gpu_launches']/150:.0f} clocks {d['clocks']['sm_mhz']}")
PY
}
run "merged heads" X=1
run "MPN_MERGE_HEADS=0" MPN_MERGE_HEADS=0
run "merged heads (rep)" X=1
