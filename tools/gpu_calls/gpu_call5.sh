#!/bin/bash
mkdir -p gpurun_out
RNF_NVCC_EXTRA="-DRNF_TC_TRACE=1" timeout 300 python tools/tc_timeline.py > gpurun_out/r02_timeline_t4.txt 2>&1; echo rc=$?
cat gpurun_out/r02_timeline_t4.txt | cut -c1-400
