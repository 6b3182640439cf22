#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 600 -k "inverse or full_size or dedup or drop_in" > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest11.log; grep "inv  rows" gpurun_out/r02_pytest11.log | grep "/tc\]" | head -20
timeout 300 python tools/dropin_probe.py 2>&1 | grep -v "total layers" | tee gpurun_out/r02_dropin_probe.txt
PREFIX=x_ STEPS=5 bash tools/ab2.sh 2>&1 | tee gpurun_out/r02_ab11.log
