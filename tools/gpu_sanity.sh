#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1; echo "all gpu tests exit $?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/all_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
