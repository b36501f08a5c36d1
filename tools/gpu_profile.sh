#!/bin/bash
# GPU box: bench lines + smoke + ncu launch list + ncu full captures of the top kernels.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -c 600 gpurun_out/bench_ref.json
# launch list (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 120 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 34 -c 4 -o gpurun_out/prof_tc \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_pool_fused -s 2 -c 1 -o gpurun_out/prof_roi \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_roi.log 2>&1; echo "ncu roi exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_ -s 8 -c 3 -o gpurun_out/prof_nms \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_nms.log 2>&1; echo "ncu nms exit $?"
ls -la gpurun_out
