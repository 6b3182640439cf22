#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s --timeout 600 > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest9.log; grep "full_grid\|config2\|config3" gpurun_out/r02_pytest9.log | head
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
timeout 300 python tools/dropin_probe.py 2>&1 | tail -4 | tee gpurun_out/r02_dropin_probe.txt
PREFIX=x_ STEPS=5 bash tools/ab2.sh 2>&1 | tee gpurun_out/r02_ab9.log
