#!/bin/bash
mkdir -p gpurun_out
python tools/timeline.py 2>&1 | tee gpurun_out/timeline.log
MPN_TC_PDL=0 python tools/timeline.py 2>&1 | tail -22 > gpurun_out/timeline_nopdl.log; tail -3 gpurun_out/timeline_nopdl.log
