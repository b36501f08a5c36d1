#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:maxpyr_all -c 1 -f -o gpurun_out/prof_pyr python tools/nms_diag.py > gpurun_out/ncu_pyr.log 2>&1; echo "ncu pyr exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nms_scan_warp_kernel -s 2 -c 1 -f -o gpurun_out/prof_nms4 python tools/nms_diag.py > gpurun_out/ncu_nms4.log 2>&1; echo "ncu nms exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv1_tc -c 1 -f -o gpurun_out/prof_c1tc python tools/nms_diag.py > gpurun_out/ncu_c1tc.log 2>&1; echo "ncu c1 exit $?"
tail -3 gpurun_out/ncu_pyr.log
