#!/bin/bash
# compute-sanitizer passes over the smoke (forward + inverse + grid, all kernels of the tc path): memcheck, synccheck, racecheck
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke" gpurun_out/r02_sanitizer_$tool.log | tail -4
done
