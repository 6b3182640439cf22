#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -m gpu -x -q -s --timeout 600 -k "train or gradients or composed or segment or autograd or smith36c" > gpurun_out/r02_pytest15.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest15.log
for rep in 1 2; do PREFIX=x_ STEPS=8 bash tools/ab2.sh 2>&1 | tee -a gpurun_out/r02_ab15.log; done
