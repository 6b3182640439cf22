#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 600 > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest14.log; grep "worst relative gradient" gpurun_out/r02_pytest14.log
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
MODE=tc timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4_cfg2 -f python tools/profile_grid.py > gpurun_out/r02_prof_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_row -s 1 -c 1 -o gpurun_out/r02_inv -f python tools/profile_inverse.py > gpurun_out/r02_prof_inv.log 2>&1; echo "ncu inv rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
