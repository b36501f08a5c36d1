#!/bin/bash
# round 2, final single-GPU validation of the tree as committed: the whole GPU suite (parity log on), smoke, both bench arms as the driver runs them
mkdir -p gpurun_out; S=gpurun_out/summary_z.txt; : > $S
export MPN_PARITY_LOG=gpurun_out/r02z_parity_errors.jsonl; : > $MPN_PARITY_LOG
timeout 2400 python -m pytest tests -q -m gpu -rxXs -p no:cacheprovider --durations=5 > gpurun_out/z_all_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a $S; tail -8 gpurun_out/z_all_gpu_tests.log | tee -a $S
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $S
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err; echo "bench reference exit $?" | tee -a $S
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/z_bench_n1_s20.json 2> gpurun_out/z_bench_n1_s20.err; echo "bench n1 s20 exit $?" | tee -a $S
python bench.py > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err; echo "bench n1 exit $?" | tee -a $S
python - <<'PY' 2>&1 | tee -a $S
import json
for c in ('n1_s20', 'n1'):
    d = json.load(open(f'gpurun_out/z_bench_{c}.json')); r = d['roofline']
    print(f"{c:8s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} p50 {d['ms_per_image_p50']:.4f} e2e {d['e2e']['value']:9.0f} raw {d['e2e_raw']['value']:9.0f} frac {r['frac']:.3f} issued {r['issued_frac']:.3f} roi {r['roi_pool']['frac']:.3f} launches {d['gpu_launches']} clk {d['clocks']}")
    print('   cpu_baseline', d.get('cpu_baseline'))
d = json.load(open('gpurun_out/z_bench_reference.json')); print('reference', d['value'], d['steps_timed'], d['cpu_baseline']['cores'])
a = json.load(open('gpurun_out/z_bench_n1_s20.json'))['config']; print('same config:', a == d['config'])
PY
cat $MPN_PARITY_LOG | tail -5 | tee -a $S
