#!/bin/bash
# round 2, call E (EIGHT B200s, gpurun --gpus 8): the 1 -> 8 curve as the driver runs it (--steps 20 --warmup 5), BASELINE configs[2..4] at 8 GPUs
mkdir -p gpurun_out; : > gpurun_out/summary_e.txt
run() { n=$1; name=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/bench_e_$name.json 2> gpurun_out/bench_e_$name.err
  else NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" \
         > gpurun_out/bench_e_$name.json 2> gpurun_out/bench_e_$name.err; fi
  echo "bench $name exit $?" | tee -a gpurun_out/summary_e.txt; }
nvidia-smi -L | tee -a gpurun_out/summary_e.txt
run 1 n1 --steps 20 --warmup 5 --no-cpu-baseline
run 8 n8 --steps 20 --warmup 5
run 4 n4 --steps 20 --warmup 5
run 8 n8_s200 --steps 200 --warmup 5
run 8 resnet50_n8 --config resnet50 --steps 20 --warmup 5
run 8 mpn_n8 --config multipathnet --steps 20 --warmup 5
run 8 nms_n8 --config nms_sweep --no-cpu-baseline
run 8 n8_reference --impl reference --steps 2 --warmup 1
python - <<'PY' 2>&1 | tee -a gpurun_out/summary_e.txt
import json
v = {}
for c in ('n1', 'n4', 'n8', 'n8_s200', 'resnet50_n8', 'mpn_n8'):
    try:
        d = json.loads(open(f'gpurun_out/bench_e_{c}.json').read().strip().splitlines()[-1])
        v[c] = d['value']
        print(f"{c:12s} value {d['value']:10.0f} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:10.0f} collective ms {d['collective']['ms']:.4f} per-rank min/max {min(d['per_rank_loop_ms_per_step']):.4f}/{max(d['per_rank_loop_ms_per_step']):.4f}")
    except Exception as e:
        print(c, 'ERR', e)
for n in (4, 8):
    if 'n1' in v and f'n{n}' in v: print(f'efficiency N={n}: {v[f"n{n}"] / (n * v["n1"]):.4f}')
for c in ('nms_n8',):
    try:
        d = json.loads(open(f'gpurun_out/bench_e_{c}.json').read().strip().splitlines()[-1])
        print(c, {k: round(x['ms_per_image'], 3) for k, x in d['sweep'].items()})
    except Exception as e:
        print(c, 'ERR', e)
PY
