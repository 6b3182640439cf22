#!/bin/bash
mkdir -p gpurun_out
for m in tc tc_x2; do
MODE=$m timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_$m -f python tools/profile_grid.py > gpurun_out/r02_prof_$m.log 2>&1; echo "ncu $m rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -3
