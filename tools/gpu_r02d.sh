#!/bin/bash
# round 2, call D (TWO B200s, gpurun --gpus 2): the library-issued all-gather on real NCCL, the N=2 bench lines, the class-sharded NMS sweep at N=2,
# and a re-run of the tests changed since call F (parity figures logged to r02_parity_errors.jsonl)
mkdir -p gpurun_out; S=gpurun_out/summary_d.txt; : > $S
export MPN_PARITY_LOG=gpurun_out/r02_parity_errors.jsonl; : > $MPN_PARITY_LOG
nvidia-smi -L | tee -a $S
timeout 900 python -m pytest tests/test_dist_gpu.py -q -m gpu -x -rs -p no:cacheprovider > gpurun_out/d_tests_dist.log 2>&1
echo "dist gpu test exit $?" | tee -a $S; tail -3 gpurun_out/d_tests_dist.log | tee -a $S
timeout 1200 python -m pytest tests/test_engine_gpu.py tests/test_roi_product_gpu.py tests/test_model_gpu.py -q -m gpu -rs -p no:cacheprovider > gpurun_out/d_tests_changed.log 2>&1
echo "changed tests exit $?" | tee -a $S; tail -4 gpurun_out/d_tests_changed.log | tee -a $S
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench_n1.json 2> gpurun_out/d_bench_n1.err; echo "bench n1 exit $?" | tee -a $S
tr() { n=$1; name=$2; shift 2
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" \
     > gpurun_out/d_bench_$name.json 2> gpurun_out/d_bench_$name.err; echo "bench $name exit $?" | tee -a $S; }
tr 2 n2 --steps 20 --warmup 5
tr 2 n2_s200 --steps 200 --warmup 5
tr 2 n2_reference --impl reference --steps 2 --warmup 1
tr 2 nms_n2 --config nms_sweep --no-cpu-baseline
tr 2 mpn_n2 --config multipathnet --steps 20 --warmup 5
python - <<'PY' 2>&1 | tee -a gpurun_out/summary_d.txt
import json
v = {}
for c in ('n1', 'n2', 'n2_s200', 'mpn_n2'):
    try:
        d = json.loads(open(f'gpurun_out/d_bench_{c}.json').read().strip().splitlines()[-1])
        v[c] = d['value']
        print(f"{c:10s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:9.0f} collective ms {d['collective']['ms']:.4f} per-rank {['%.4f' % x for x in d['per_rank_loop_ms_per_step']]} dets/img {d['collective']['detections_per_image_mean']:.1f}")
    except Exception as e:
        print(c, 'ERR', e)
if 'n1' in v and 'n2' in v: print('efficiency N=2 (steps 20):', v['n2'] / (2 * v['n1']))
try:
    d = json.loads(open('gpurun_out/d_bench_nms_n2.json').read().strip().splitlines()[-1])
    print('nms_n2', {k: round(x['ms_per_image'], 3) for k, x in d['sweep'].items()})
except Exception as e: print('nms_n2 ERR', e)
PY
cat $MPN_PARITY_LOG | tee -a $S
