#!/bin/bash
# round-2 GPU call 1: parity of the half-angle mixture + A/B of mbarrier wait hints
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/r02_pytest1.log
STEPS=5 timeout 900 bash tools/ab.sh 2>&1 | tee gpurun_out/r02_ab1.log
