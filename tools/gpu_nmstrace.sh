#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_nms_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
MPN_NMS_TRACE=1 python tools/nms_diag.py 2>&1 | grep -v "^class" | sort | uniq -c | sort -rn | head -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms_ -s 12 -c 3 python tools/nms_diag.py 2>&1 | grep -E "gpu__time" | head -4
python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),{k:round(v,4) for k,v in d['roofline']['by_category_ms_per_step'].items()})"
