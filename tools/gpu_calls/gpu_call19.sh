#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 600 -k "s_clu" > gpurun_out/r02_pytest19.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest19.log; grep "s_clu" gpurun_out/r02_pytest19.log | head; grep -E "^E  " gpurun_out/r02_pytest19.log | head -8
