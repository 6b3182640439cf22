#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke28.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke28.log
if [ $rc -ne 0 ]; then exit 1; fi
PREFIX=w_ bash tools/ab_ncu.sh 2>&1 | tee gpurun_out/ab_ncu_w.txt
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product.so; cp tools/_build/t_trace.so rotationnormflow_b200/librnf_b200.so
timeout 120 python tools/tc_timeline.py > gpurun_out/r02_t4_timeline_service.txt 2>&1; tail -16 gpurun_out/r02_t4_timeline_service.txt | cut -c1-330
cp tools/_build/.product.so rotationnormflow_b200/librnf_b200.so
