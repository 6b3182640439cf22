#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do PREFIX=x_ STEPS=8 bash tools/ab2.sh 2>&1 | tee -a gpurun_out/r02_ab17.log; done
cp tools/_build/x_1epiasync.so rotationnormflow_b200/librnf_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "forward_parity and tc" 2>&1 | tail -2
