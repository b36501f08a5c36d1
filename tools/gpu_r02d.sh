#!/bin/bash
# round 2, call D (TWO B200s, gpurun --gpus 2): the library-issued all-gather on real NCCL + the N=2 bench line + sanitizer logs
mkdir -p gpurun_out; : > gpurun_out/summary_d.txt
nvidia-smi -L | tee -a gpurun_out/summary_d.txt
timeout 900 python -m pytest tests/test_dist_gpu.py -q -m gpu -x -rs -p no:cacheprovider > gpurun_out/tests_dist.log 2>&1
echo "dist gpu test exit $?" | tee -a gpurun_out/summary_d.txt; tail -5 gpurun_out/tests_dist.log | tee -a gpurun_out/summary_d.txt
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_d_n1.json 2> gpurun_out/bench_d_n1.err; echo "bench n1 exit $?" | tee -a gpurun_out/summary_d.txt
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 \
   > gpurun_out/bench_d_n2.json 2> gpurun_out/bench_d_n2.err; echo "bench n2 exit $?" | tee -a gpurun_out/summary_d.txt
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 200 --warmup 5 \
   > gpurun_out/bench_d_n2_s200.json 2> gpurun_out/bench_d_n2_s200.err; echo "bench n2 s200 exit $?" | tee -a gpurun_out/summary_d.txt
python - <<'PY' 2>&1 | tee -a gpurun_out/summary_d.txt
import json
v = {}
for c in ('d_n1', 'd_n2', 'd_n2_s200'):
    try:
        d = json.loads(open(f'gpurun_out/bench_{c}.json').read().strip().splitlines()[-1])
        v[c] = d['value']
        print(f"{c:10s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:9.0f} collective ms {d['collective']['ms']:.4f} per-rank {['%.4f' % x for x in d['per_rank_loop_ms_per_step']]} dets/img {d['collective']['detections_per_image_mean']:.1f}")
    except Exception as e:
        print(c, 'ERR', e)
if 'd_n1' in v and 'd_n2' in v: print('efficiency N=2 (steps 20):', v['d_n2'] / (2 * v['d_n1']))
PY
# compute-sanitizer on the smoke path (tcgen05 conv / GEMM kernels, cluster ROI kernel, NMS, pack) and the small op tests
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1
echo "memcheck smoke exit $?" | tee -a gpurun_out/summary_d.txt; tail -4 gpurun_out/r02_sanitizer_memcheck_smoke.log | tee -a gpurun_out/summary_d.txt
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1
echo "racecheck smoke exit $?" | tee -a gpurun_out/summary_d.txt; tail -4 gpurun_out/r02_sanitizer_racecheck_smoke.log | tee -a gpurun_out/summary_d.txt
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_post_gpu.py tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "not full_size" > gpurun_out/r02_sanitizer_memcheck_ops.log 2>&1
echo "memcheck ops exit $?" | tee -a gpurun_out/summary_d.txt; tail -4 gpurun_out/r02_sanitizer_memcheck_ops.log | tee -a gpurun_out/summary_d.txt
