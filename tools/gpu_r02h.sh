#!/bin/bash
# round 2, call H (TWO B200s): model replicas per GPU — the new test file, bench at K = 1 / 2 / 3 on one GPU, the torchrun path at N = 2
mkdir -p gpurun_out; S=gpurun_out/summary_h.txt; : > $S
timeout 900 python -m pytest tests/test_replicas_gpu.py tests/test_dist_gpu.py tests/test_post_gpu.py -q -m gpu -rs -p no:cacheprovider > gpurun_out/h_tests.log 2>&1
echo "tests exit $?" | tee -a $S; tail -4 gpurun_out/h_tests.log | tee -a $S
run() { name=$1; shift; python bench.py --no-cpu-baseline "$@" > gpurun_out/h_bench_$name.json 2> gpurun_out/h_bench_$name.err; echo "bench $name exit $?" | tee -a $S; tail -2 gpurun_out/h_bench_$name.err >> $S; }
run k2_s20 --steps 20 --warmup 5
run k1 --replicas 1
run k2 --replicas 2
run k3 --replicas 3
run k2_again --replicas 2
run mpn_k2 --config multipathnet --steps 60
run mpn_k1 --config multipathnet --steps 60 --replicas 1
run resnet50_k2 --config resnet50 --steps 40
tr() { n=$1; name=$2; shift 2
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" \
     > gpurun_out/h_bench_$name.json 2> gpurun_out/h_bench_$name.err; echo "bench $name exit $?" | tee -a $S; tail -2 gpurun_out/h_bench_$name.err >> $S; }
tr 2 n2 --steps 20 --warmup 5
tr 2 n2_s200 --steps 200 --warmup 5
python - <<'PY' 2>&1 | tee -a gpurun_out/summary_h.txt
import json
v = {}
for c in ('k2_s20', 'k1', 'k2', 'k3', 'k2_again', 'mpn_k2', 'mpn_k1', 'resnet50_k2', 'n2', 'n2_s200'):
    try:
        d = json.loads(open(f'gpurun_out/h_bench_{c}.json').read().strip().splitlines()[-1]); v[c] = d['value']
        print(f"{c:12s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} p50 {d['ms_per_image_p50']:.4f} e2e {d['e2e']['value']:9.0f} sync {d['e2e']['sync_value']:9.0f} raw {(d.get('e2e_raw') or {}).get('value', 0):9.0f} coll ms {d['collective']['ms']:.4f} launches {d['gpu_launches']} clk {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(c, 'ERR', e)
if 'k2_s20' in v and 'n2' in v: print('efficiency N=2 (steps 20):', v['n2'] / (2 * v['k2_s20']))
PY
