#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke41.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke41.log
if [ $rc -ne 0 ]; then exit 1; fi
PREFIX=c_ TMO=60 bash tools/ab_ncu.sh 2>&1 | tee gpurun_out/ab_ncu_clean.txt
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest41.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest41.log
