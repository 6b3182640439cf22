#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_row -s 1 -c 1 -o gpurun_out/r02_inv2 -f python tools/profile_inverse.py > gpurun_out/r02_prof_inv2.log 2>&1; echo "ncu inv rc=$?"
MODE=tc timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4_cfg2b -f python tools/profile_grid.py > gpurun_out/r02_prof_cfg2b.log 2>&1; echo "ncu cfg2 rc=$?"
ls -la gpurun_out/r02_inv2.ncu-rep gpurun_out/r02_t4_cfg2b.ncu-rep
