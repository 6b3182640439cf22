#!/bin/bash
mkdir -p gpurun_out
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product2.so
for v in tools/_build/h_*.so; do cp "$v" rotationnormflow_b200/librnf_b200.so; echo "== $v"; timeout 300 python tools/inverse_hard_probe.py 2>&1 | tail -4; done | tee gpurun_out/r02_inverse_hard_probe.txt
cp tools/_build/.product2.so rotationnormflow_b200/librnf_b200.so
