#!/bin/bash
# run with: gpurun --gpus 2 -- bash tools/gpu_2gpu.sh
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 100 --warmup 5 \
   > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"
tail -c 2500 gpurun_out/bench_n2.json; tail -n 5 gpurun_out/bench_n2.err
python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1b.err; echo "bench n1 exit $?"
python - <<'PY'
import json
for c in ('n2','n1_same_box'):
    try:
        d=json.loads(open(f'gpurun_out/bench_{c}.json').read().strip().splitlines()[-1])
        print(c,'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3), d.get('clocks'))
    except Exception as e: print(c,'ERR',e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 \
   > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 exit $?"; tail -c 800 gpurun_out/bench_ref_n2.json
