#!/bin/bash
# ncu capture + phase timeline of the service-warp kernel (half-epilogue build)
mkdir -p gpurun_out
MODE=tc timeout 400 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4svc_cfg2 -f python tools/profile_grid.py > gpurun_out/r02_prof_svc.log 2>&1; echo "ncu rc=$?"
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product.so; cp tools/_build/t_trace.so rotationnormflow_b200/librnf_b200.so
timeout 120 python tools/tc_timeline.py > gpurun_out/r02_t4_timeline_service.txt 2>&1; tail -16 gpurun_out/r02_t4_timeline_service.txt | cut -c1-330
cp tools/_build/.product.so rotationnormflow_b200/librnf_b200.so
