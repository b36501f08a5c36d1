#!/bin/bash
# Runs on the GPU box: bring-up diagnostics, then each GPU test file in its own process (bounded).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$1" != "notl" ]; then timeout 900 python tools/first_light.py > gpurun_out/first_light.stdout 2>&1; fi
for f in test_nms_gpu test_ops_gpu test_engine_gpu test_model_gpu; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -3 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
