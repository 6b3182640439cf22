#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "inverse or full_size" 2>&1 | tail -2
NGPU=4 bash tools/gpu_calls/gpu_call16.sh 2>&1 | grep -E "^N |rc="
