#!/bin/bash
# round 2, final multi-GPU run (EIGHT B200s): N = 1 / 4 / 8 as the driver runs them (--steps 20 --warmup 5), two replicas per GPU
mkdir -p gpurun_out; S=gpurun_out/summary_y.txt; : > $S
run() { n=$1; name=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/y_bench_$name.json 2> gpurun_out/y_bench_$name.err
  else NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" \
         > gpurun_out/y_bench_$name.json 2> gpurun_out/y_bench_$name.err; fi
  echo "bench $name exit $?" | tee -a $S; }
run 1 n1 --steps 20 --warmup 5 --no-cpu-baseline
run 8 n8 --steps 20 --warmup 5
run 4 n4 --steps 20 --warmup 5
python - <<'PY' 2>&1 | tee -a $S
import json
v = {}
for c in ('n1', 'n4', 'n8'):
    try:
        d = json.loads(open(f'gpurun_out/y_bench_{c}.json').read().strip().splitlines()[-1]); v[c] = d['value']
        print(f"{c:4s} value {d['value']:10.0f} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:10.0f} collective ms {d['collective']['ms']:.4f} per-rank min/max {min(d['per_rank_loop_ms_per_step']):.4f}/{max(d['per_rank_loop_ms_per_step']):.4f} replicas {d['replicas_per_gpu']}")
    except Exception as e:
        print(c, 'ERR', e)
for n in (4, 8):
    if 'n1' in v and f'n{n}' in v: print(f'efficiency N={n}: {v[f"n{n}"] / (n * v["n1"]):.4f}')
PY
