#!/bin/bash
mkdir -p gpurun_out
PREFIX=v_ TMO=60 bash tools/ab_ncu.sh 2>&1 | tee gpurun_out/ab_ncu_v.txt
