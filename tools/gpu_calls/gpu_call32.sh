#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke32.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke32.log
if [ $rc -ne 0 ]; then exit 1; fi
PREFIX=v_ TMO=60 bash tools/ab_ncu.sh 2>&1 | tee gpurun_out/ab_ncu_named.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "forward or grid or tiles_in_flight" 2>&1 | tail -2
