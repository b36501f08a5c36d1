#!/bin/bash
# same box, back to back: contribution of each engine feature to the default workload (value = device-resident proposals/s)
mkdir -p gpurun_out; : > gpurun_out/ablation.txt
run() {  # name, env...
  name=$1; shift
  out=$(env "$@" python bench.py --steps 150 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$name" "$out" <<'PY' | tee -a gpurun_out/ablation.txt
import json,sys
d=json.loads(sys.argv[2])
print(f"{sys.argv[1]:28s} value {d['value']:9.0f}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:9.0f}  sync {d['e2e']['sync_value']:9.0f}  clocks {d['clocks']['sm_mhz']}")
PY
}
run "all on (default)" X=1
run "MPN_CONV1_TC=0" MPN_CONV1_TC=0
run "MPN_TC_PDL=0" MPN_TC_PDL=0
run "MPN_TC_STREAMK=0" MPN_TC_STREAMK=0
run "MPN_TC_FUSE_POOL=0" MPN_TC_FUSE_POOL=0
run "MPN_TC_R3=0" MPN_TC_R3=0
run "MPN_TC_CTA_GROUP=1" MPN_TC_CTA_GROUP=1
run "all on (repeat)" X=1
