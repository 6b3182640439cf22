#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/active_tiles_probe.py 2>&1 | tee gpurun_out/r02_active_tiles.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "forward_parity or edge or short_and_odd or grid_log_prob" 2>&1 | tail -2
