#!/bin/bash
# round 2, call G (one B200): experiments — replicas per GPU (tools/two_replicas.py), conv5 on the generic kernel (MPN_TC_R3_MINPIX),
# ncu --set full of the w16 fc6 launch, and the whole GPU suite once more with the parity log on
mkdir -p gpurun_out; S=gpurun_out/summary_g.txt; : > $S
export MPN_PARITY_LOG=gpurun_out/r02g_parity_errors.jsonl; : > $MPN_PARITY_LOG
timeout 600 python tools/two_replicas.py vgg16_frcnn 300 2>&1 | grep replicas | tee -a $S
timeout 600 python tools/two_replicas.py multipathnet 80 2>&1 | grep replicas | tee -a $S
run() { name=$1; cfg=$2; steps=$3; shift 3
  env "$@" python bench.py --config $cfg --steps $steps --no-cpu-baseline > gpurun_out/g_bench_$name.json 2> gpurun_out/g_bench_$name.err; echo "bench $name exit $?" >> $S; }
run n1 vgg16_frcnn 200 X=1
run n1_minpix vgg16_frcnn 200 MPN_TC_R3_MINPIX=2000
run n1_again vgg16_frcnn 200 X=1
run mpn multipathnet 60 X=1
python - <<'PY' 2>&1 | tee -a $S
import json
for c in ('n1', 'n1_minpix', 'n1_again', 'mpn'):
    try:
        d = json.load(open(f'gpurun_out/g_bench_{c}.json'))
        r = d['roofline']; b = r['by_category_ms_per_step']
        print(f"{c:12s} value {d['value']:9.0f} ms/step {d['ms_per_step']:.4f} e2e {d['e2e']['value']:9.0f} tc {b['conv_gemm_tc']:.4f} roi {b['roi_pool']:.4f} nms {b['nms']:.4f} tc frac {r['frac']:.3f} clk {d['clocks']['sm_mhz']}")
    except Exception as e:
        print(c, 'ERR', e)
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_tc -s 4 -c 1 -f -o gpurun_out/r02g_fc6_w16 python tools/prof_step.py vgg16_frcnn 3 > gpurun_out/ncu_r02g_fc6.log 2>&1; echo "ncu fc6 exit $?" | tee -a $S
timeout 2400 python -m pytest tests -q -m gpu -rxXs -p no:cacheprovider --durations=5 > gpurun_out/g_all_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a $S; tail -8 gpurun_out/g_all_gpu_tests.log | tee -a $S
