#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 600 -k "inverse or full_size or drop_in" > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest20.log; grep "inv  rows" gpurun_out/r02_pytest20.log | grep "/tc\]" | tail -6
for rep in 1 2; do PREFIX=x_ STEPS=4 bash tools/ab2.sh 2>&1 | tee -a gpurun_out/r02_ab20.log; done
