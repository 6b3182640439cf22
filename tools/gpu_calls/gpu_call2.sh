#!/bin/bash
mkdir -p gpurun_out
RNF_TEST_MODES=tc_x2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 60 -k "forward_parity or short_and_odd or grid_log_prob or spread or edge" > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest2.log
MODE=tc_x2 PREFIX=x_ bash tools/ab2.sh 2>&1 | tee gpurun_out/r02_ab2.log
