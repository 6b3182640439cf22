#!/bin/bash
# service-warp kernel: quick smoke (short timeout: a hang must not eat the budget), parity, then A/B against the old kernel
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke25.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke25.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "forward or grid" 2>&1 | tail -3
for rep in 1 2; do PREFIX=x_ STEPS=5 TMO=100 bash tools/ab2.sh; done 2>&1 | tee gpurun_out/ab_service.txt
