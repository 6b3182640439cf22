#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 > gpurun_out/r02_pytest22.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest22.log; grep -E "^E  |^tests.*Error|FAILED" gpurun_out/r02_pytest22.log | head -12
